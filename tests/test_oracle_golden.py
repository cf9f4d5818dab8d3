"""Pin the numpy oracle (oracle/real3d_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from oracle import real3d_oracle as orc
from pcl_augmentation_b200 import synth
from tests.helpers import case_from_golden, load_golden, parse_inserted

FN_IMG = (64, 512)


def _fn_projection_inputs():
    shape = synth.ScanShape(32, 300, 2.0, -24.8)
    pcl, labels = synth.make_scan(5, shape)
    pcl5 = np.hstack((pcl, labels.reshape(-1, 1))).astype(np.float64)
    obj = synth.make_cut_object(77, "Cyclist", False, shape, (6.0, 9.0))["pcl"]
    obj[:, 2] += np.linspace(-0.5, 3.0, len(obj))
    return pcl5, obj


def test_projection_and_closing_match_reference():
    g = load_golden("fn_projection")
    pcl5, obj = _fn_projection_inputs()
    assert synth.array_digest(pcl5, obj) == str(g["in_digest"])
    pc = orc.add_space_for_spherical(pcl5)
    pc, mx, mn = orc.fill_spherical(pc)
    assert mx == float(g["max_el"]) and mn == float(g["min_el"])
    np.testing.assert_array_equal(pc[:, 3:6], g["sph"])
    train, label, pc = orc.geometrical_front_view(pc, *FN_IMG, mx, mn)
    np.testing.assert_array_equal(pc[:, 8].astype(np.int32), g["pix"])
    np.testing.assert_array_equal(train, g["train"])
    np.testing.assert_array_equal(label.astype(np.int8), g["label"])
    s_train, s_label = orc.smooth_out(train, label)
    np.testing.assert_array_equal(s_label.astype(np.int8), g["s_label"])
    np.testing.assert_array_equal(s_train, g["s_train"])          # bit-exact fp64 neighbour means
    # object projected with the scene's elevation range, sample=True row skipping
    opc = orc.add_space_for_spherical(obj)
    opc, _, _ = orc.fill_spherical(opc)
    o_train, o_label, opc = orc.geometrical_front_view(opc, *FN_IMG, mx, mn, sample=True)
    np.testing.assert_array_equal(opc[:, 8].astype(np.int32), g["o_pix"])
    assert (g["o_pix"] < 0).sum() > 0
    np.testing.assert_array_equal(o_train, g["o_train"])
    os_train, os_label = orc.smooth_out(o_train, o_label)
    np.testing.assert_array_equal(os_train, g["os_train"])
    np.testing.assert_array_equal(os_label.astype(np.int8), g["os_label"])


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _fn_projection_full_inputs():
    pcl, labels = synth.make_scan(6, synth.KITTI_SHAPE)
    return np.hstack((pcl, labels.reshape(-1, 1))).astype(np.float64)


def check_projection_full(g, pc, mx, mn, train, label, s_train, s_label, exact_elevation=True):
    """A1-A4 results on the 120 000-point scan vs what the unmodified reference produced at 112 x 1440."""
    if exact_elevation:
        assert mx == float(g["max_el"]) and mn == float(g["min_el"])
    assert _sha(pc[:, 3]) == str(g["r_sha"])
    np.testing.assert_array_equal(pc[:, 8].astype(np.int32), g["pix"])
    assert int((label == 1).sum()) == int(g["occupied"])
    assert _sha(train) == str(g["train_sha"]) and _sha(label.astype(np.int8)) == str(g["label_sha"])
    filled = np.flatnonzero((s_label == 1) & (label != 1)).astype(np.int32)
    np.testing.assert_array_equal(filled, g["filled"])
    np.testing.assert_array_equal(s_train.ravel()[filled], g["filled_values"])      # bit-exact fp64 neighbour means
    assert _sha(s_train) == str(g["s_train_sha"]) and _sha(s_label.astype(np.int8)) == str(g["s_label_sha"])


def test_projection_and_closing_match_reference_at_baseline_size():
    g = load_golden("fn_projection_full")
    pcl5 = _fn_projection_full_inputs()
    assert synth.array_digest(pcl5) == str(g["in_digest"])
    pc = orc.add_space_for_spherical(pcl5)
    pc, mx, mn = orc.fill_spherical(pc)
    train, label, pc = orc.geometrical_front_view(pc, 112, 1440, mx, mn)
    s_train, s_label = orc.smooth_out(train, label)
    check_projection_full(g, pc, mx, mn, train, label, s_train, s_label)


def test_cut_bounding_box_matches_reference():
    g = load_golden("fn_cut_bbox")
    rng = np.random.default_rng(99)
    pts = rng.uniform(-6, 6, (6000, 5))
    pts[:, 2] = rng.uniform(-2, 3, 6000)
    pts = pts.astype(np.float32).astype(np.float64)
    assert synth.array_digest(pts) == str(g["in_digest"])
    for b, packed in zip(g["boxes"], g["masks"]):
        anno = {"center": {"x": b[0], "y": b[1], "z": b[2]}, "rotation": {"x": b[3], "y": b[4], "z": b[5], "w": b[6]},
                "length": b[7], "width": b[8], "height": b[9], "class": "Car"}
        m = orc.cut_bounding_box_mask(pts, anno)
        np.testing.assert_array_equal(m, np.unpackbits(packed)[:len(pts)].astype(bool))


def _scene9(case):
    pcl5 = case.pcl5.copy()
    if case.task == "od":
        pcl5[pcl5[:, 4] != case.config["labels"]["Road"], 4] = 1
    pc = orc.add_space_for_spherical(pcl5)
    pc, mx, mn = orc.fill_spherical(pc)
    _, _, pc = orc.geometrical_front_view(pc, 112, 1440, mx, mn)
    return pcl5, pc


@pytest.mark.parametrize("task", ["od", "ss"])
@pytest.mark.parametrize("mode,fast", [("cumulative", False), ("cumulative", True), ("closed", True)])
def test_find_possible_places_matches_reference(task, mode, fast):
    g = load_golden(f"fn_places_{task}")
    spec, case = case_from_golden(g)
    original, scene = _scene9(case)
    classes = case.config["insertion"]["classes"]
    read = orc.read_label_line_ss if task == "ss" else orc.read_label_line_od
    annos = [read(l) for l in case.box_lines]
    if task == "ss":
        map_arr, map_move = orc.addjust_map_2(orc.FreshDict(case.map_data), scene, case.pose)
        np.testing.assert_array_equal(np.argwhere(map_arr == 4).astype(np.int32), g["map_adjusted_cells"])
    total = 0
    for ci, cls in enumerate(classes):
        for j in range(4):
            name, sample = case.db[cls][j]
            sd = orc.FreshDict(sample)
            if task == "od":
                placement = case.config["insertion"]["placement"][cls]
                pcls, ans, rots = orc.find_possible_places_od(scene, annos, sd, case.maps[placement], original,
                                                              case.config, mode=mode, fast=fast)
            else:
                pcls, ans, rots = orc.find_possible_places_ss(scene, annos, sd, map_arr, map_move, original,
                                                              case.pose, case.config, mode=mode, fast=fast)
            key = f"c{ci}_s{j}"
            np.testing.assert_array_equal(np.array(rots, dtype=np.int32), g[key + "_rots"])
            total += len(rots)
            if rots:
                for n, p in enumerate(g[key + "_pick"]):
                    np.testing.assert_allclose(pcls[p][:, :3], g[key + "_xyz"][n], rtol=0, atol=1e-9)
                    box = g[key + "_box"][n]
                    c = ans[p]["center"]
                    np.testing.assert_allclose([c["x"], c["y"], c["z"]], box[:3], rtol=0, atol=1e-9)
                    m_ref = orc.R.from_quat(box[3:7]).as_matrix()
                    np.testing.assert_allclose(orc.box_matrix(ans[p]), m_ref, rtol=0, atol=1e-9)
    assert total > 100


E2E = ["e2e_od_a", "e2e_od_b", "e2e_od_c", "e2e_ss_a", "e2e_ss_b", "e2e_ss_c",
       # BASELINE-size scans (120 000 / 124 992 points) and the very large, very close cut objects (> 4096 points)
       "e2e_od_full", "e2e_ss_full", "e2e_ss_big"]


def run_oracle_e2e(g, case, mode="cumulative", fast=True):
    pose = g["used_pose"] if "used_pose" in g.files else case.pose
    return orc.augment_scan(case.task, case.pcl5, case.box_lines, case.db, case.schedule.counts,
                            case.schedule.perms, case.config, maps=case.maps, map_data=case.map_data,
                            transform_matrix=pose, mode=mode, fast=fast)


def check_e2e_against_golden(g, case, res, exact_bytes=True):
    """Compare an augment_scan-style result dict with what the reference's insertion.py wrote."""
    task = case.task
    want = parse_inserted(str(g["inserted"]))
    got = [(n, int(r)) for (n, r, _c) in res["inserted"]]
    assert got == want
    if int(g["n_out"]) < 0:
        assert not res["inserted"]
        return
    out = orc.save_arrays(task, res)
    n0 = len(case.pcl5)
    keep = np.unpackbits(g["keep_orig"])[:n0].astype(bool)
    np.testing.assert_array_equal(res["keep_orig"], keep)
    assert len(out["velodyne"]) == int(g["n_out"])
    n_kept = int(g["n_kept"])
    np.testing.assert_array_equal(out["velodyne"][:n_kept], case.pcl5[keep][:, :4].astype(np.float32))
    np.testing.assert_allclose(out["velodyne"][n_kept:], g["tail"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["check"], g["check"], rtol=0, atol=1e-6)
    if task == "ss":
        np.testing.assert_array_equal(out["labels"][n_kept:].ravel(), g["tail_labels"])
    if exact_bytes:
        import hashlib
        assert hashlib.sha256(np.ascontiguousarray(out["velodyne"]).tobytes()).hexdigest() == str(g["velodyne_sha"])
    if task == "od":
        ref_lines = str(g["label_2"]).splitlines(keepends=True)
        assert ref_lines[len(case.box_lines):] == res["lines"]


@pytest.mark.parametrize("name", E2E)
def test_e2e_oracle_matches_reference_run(name):
    g = load_golden(name)
    spec, case = case_from_golden(g)
    res = run_oracle_e2e(g, case)
    assert res["n_events"] == int(g["n_events"])
    check_e2e_against_golden(g, case, res, exact_bytes=False)


@pytest.mark.parametrize("name", ["e2e_od_a", "e2e_ss_a"])
def test_e2e_closed_form_equals_cumulative(name):
    g = load_golden(name)
    spec, case = case_from_golden(g)
    res = run_oracle_e2e(g, case, mode="closed")
    check_e2e_against_golden(g, case, res, exact_bytes=False)


def test_rich_map_od_matches_reference_script():
    """rich-map oracle vs what the unmodified single_drivable_area_map.py wrote."""
    import json
    from oracle import rich_map_oracle as rmo
    from tests.helpers import case_from_spec
    g = load_golden("rich_map_od")
    meta = json.loads(str(g["meta"]))
    for i, (seed, n_cars) in enumerate(zip(meta["seeds"], meta["n_cars"])):
        case = case_from_spec(dict(task="od", seed=seed, counts=[1, 1], n_cars=n_cars))
        assert synth.case_digest(case) == str(g[f"digest{i}"])
        pcl5 = np.hstack((case.pcl5[:, :4].astype(np.float32), case.pcl5[:, 4:5]))       # KITTI.__getitem__ (od/ds:62-66)
        road, ped, min_x, min_y = rmo.rich_map_od(pcl5, case.config["labels"]["Road"])
        for key, got in (("road", road), ("ped", ped)):
            shape = tuple(g[f"{key}{i}_shape"])
            want = np.unpackbits(g[f"{key}{i}_bits"])[:shape[0] * shape[1]].reshape(shape)
            assert got.shape == shape and got.dtype == np.uint8 == np.dtype(str(g[f"{key}{i}_dtype"]))
            np.testing.assert_array_equal(got, want)
            assert [min_x, min_y] == list(g[f"{key}{i}_min"])


def test_rich_map_ss_matches_reference_script():
    """sequence-map oracle vs what the unmodified drivable_area_map.py wrote."""
    import json
    from oracle import rich_map_oracle as rmo
    from tests.helpers import rich_map_ss_cases
    g = load_golden("rich_map_ss")
    cases = rich_map_ss_cases()
    frames = []
    for i, case in enumerate(cases):
        assert synth.case_digest(case) == str(g[f"digest{i}"])
        pcl5 = np.hstack((case.pcl5[:, :4].astype(np.float32), case.pcl5[:, 4:5]))       # SemanticKITTI.__getitem__ (ss/ds:45-62)
        frames.append((pcl5, g["poses"][i]))
    labels = {int(k): v for k, v in json.loads(str(g["placement_labels"])).items()}
    got = rmo.rich_map_ss(frames, labels)
    assert got["map"].dtype == np.dtype(str(g["map_dtype"])) and got["map"].shape == g["map"].shape
    np.testing.assert_array_equal(got["map"], g["map"].astype(np.float64))
    np.testing.assert_array_equal(got["move"], g["move"])
    # the fixture exercises both ordering rules: cells hit by road AND parking points, and sticky sidewalk cells
    table = rmo.surface_classes(labels)
    world = [(T @ np.c_[p[:, :3], np.ones(len(p))].T).T for p, T in frames]
    cells = {}
    for (p, _), w in zip(frames, world):
        for lab, x, y in zip(p[:, 4], w[:, 0], w[:, 1]):
            c = table.get(int(lab))
            if c:
                cells.setdefault((int(x - got["move"][0, 0]), int(y - got["move"][1, 0])), set()).add(c)
    assert sum(1 for v in cells.values() if {1, 2} <= v) > 10
    assert sum(1 for v in cells.values() if 3 in v and len(v) > 1) > 10

