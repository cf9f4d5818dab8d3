"""GPU tier, SURVEY §8f row 3: per-frame rich-map generation (object detection) against the unmodified reference
script's output (golden) and against the oracle on full-size scans; bit-exact (uint8 cell values, extents)."""
import json
import os

import numpy as np
import pytest

from oracle import rich_map_oracle as rmo
from pcl_augmentation_b200 import synth, synth_io
from pcl_augmentation_b200.object_detection.rich_map import single_drivable_area_map as rm
from tests.helpers import case_from_spec, load_golden

pytestmark = pytest.mark.gpu


def test_rich_maps_match_reference_script_files(tmp_path):
    g = load_golden("rich_map_od")
    meta = json.loads(str(g["meta"]))
    cases = [case_from_spec(dict(task="od", seed=s, counts=[1, 1], n_cars=c)) for s, c in zip(meta["seeds"], meta["n_cars"])]
    _, _, cfg = synth_io.write_od_dataset(cases, str(tmp_path))
    for sub in ("road_maps", "pedestrian_area"):                       # the generator must write them itself
        for f in os.listdir(tmp_path / "maps" / "maps" / sub / "npz"):
            os.remove(tmp_path / "maps" / "maps" / sub / "npz" / f)
    assert rm.generate_maps(cfg, batch_size=8, log=lambda *a: None) == 2
    for i in range(2):
        for key, sub in (("road", "road_maps"), ("ped", "pedestrian_area")):
            z = np.load(tmp_path / "maps" / "maps" / sub / "npz" / f"{i:06d}.npz")
            shape = tuple(g[f"{key}{i}_shape"])
            want = np.unpackbits(g[f"{key}{i}_bits"])[:shape[0] * shape[1]].reshape(shape)
            assert z["map"].dtype == np.uint8 and z["map"].shape == shape
            np.testing.assert_array_equal(z["map"], want)
            assert [int(z["min_x"]), int(z["min_y"])] == list(g[f"{key}{i}_min"])


def test_rich_maps_full_size_batch_vs_oracle():
    scans = []
    for seed, shape in ((51, synth.KITTI_SHAPE), (52, synth.KITTI_SHAPE), (53, synth.OS128_SHAPE), (54, synth.SMALL_SHAPE)):
        pcl, labels = synth.make_scan(seed, shape, synth.make_scene_cars(seed, 6))
        scans.append((pcl, labels & 0xFFFF))
    scans.append((np.array([[0.2, -0.7, -1.7, 0.5]], dtype=np.float32), np.array([40], dtype=np.uint32)))   # one road point
    scans.append((np.array([[3.5, 2.5, -1.7, 0.5], [-4.2, 9.9, 0, 0]], dtype=np.float32), np.array([1, 1], dtype=np.uint32)))  # no road
    got = rm.drivable_area_maps_batch([s[0] for s in scans], [s[1] for s in scans], 40)
    for (pcl, labels), (road, ped) in zip(scans, got):
        w_road, w_ped, mx, my = rmo.rich_map_od(np.hstack((pcl, labels.reshape(-1, 1))).astype(np.float64), 40)
        np.testing.assert_array_equal(road["map"], w_road)
        np.testing.assert_array_equal(ped["map"], w_ped)
        assert (road["min_x"], road["min_y"], ped["min_x"], ped["min_y"]) == (mx, my, mx, my)
    single = rm.drivable_area_maps(np.hstack((scans[0][0], scans[0][1].reshape(-1, 1))), 40)
    np.testing.assert_array_equal(single[0]["map"], got[0][0]["map"])


def test_engine_runs_on_generated_maps():
    """The generated maps feed the engine exactly like the maps read from disk."""
    from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case
    case = synth.make_case("od", 4001, number_of_object=4)
    road, ped = rm.drivable_area_maps(case.pcl5, case.config["labels"]["Road"])
    inp = scan_input_from_case(case)
    inp.maps = {"Road": road, "Sidewalk": ped}
    eng = Real3DEngine("od", case.config, case.db, max_scans=1, max_points=len(case.pcl5))
    res = eng.augment_batch([inp])[0]
    eng.close()
    from oracle import real3d_oracle as orc
    ref = orc.augment_scan("od", case.pcl5, case.box_lines, case.db, case.schedule.counts, case.schedule.perms, case.config,
                           maps={"Road": road, "Sidewalk": ped}, mode="closed")
    assert [(n, int(r)) for n, r, _ in res.inserted] == [(n, int(r)) for n, r, _ in ref["inserted"]]
    assert len(res.velodyne) == len(ref["scene"])


# ------------------------------------------------------------------------------------------------ semseg (ss/rm)
def _chained_frames(seeds, shape, stride=6.5):
    from pcl_augmentation_b200.semantic_segmentation.Real3DAug.tools.datasets import SemanticKITTI
    base = synth.make_pose(seeds[0])
    frames = []
    for i, seed in enumerate(seeds):
        pcl, labels = synth.make_scan(seed, shape, synth.make_scene_cars(seed, 4))
        a = 0.2 * i
        step = np.eye(4)
        step[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        step[:3, 3] = [stride * i, 1.7 * i, 0.02 * i]
        frames.append((pcl, labels & 0xFFFF, base @ step))
    return frames


def _oracle_frames(frames):
    return [(np.hstack((p.astype(np.float64), np.asarray(l, dtype=np.float64).reshape(-1, 1))), T) for p, l, T in frames]


def test_sequence_map_matches_reference_script_file(tmp_path):
    import yaml
    from pcl_augmentation_b200.semantic_segmentation.rich_map import drivable_area_map as srm
    from tests.helpers import rich_map_ss_cases
    g = load_golden("rich_map_ss")
    cases = rich_map_ss_cases()
    _, _, cfg = synth_io.write_ss_dataset(cases, str(tmp_path))
    cfg["path"]["maps_path"] = str(tmp_path / "maps" / "small" / "npz")
    out = srm.generate_map(cfg, "00", chunk_frames=3, log=lambda *a: None)
    z = np.load(tmp_path / "maps" / "small" / "npz" / "00.npz")
    assert z["map"].dtype == np.dtype(str(g["map_dtype"])) and z["map"].shape == g["map"].shape
    np.testing.assert_array_equal(z["map"], g["map"].astype(np.float64))
    np.testing.assert_array_equal(z["move"], g["move"])
    np.testing.assert_array_equal(out["map"], z["map"])
    # frames re-read for the raster pass (nothing kept on the device) give the same map
    out2 = srm.generate_map(cfg, "00", chunk_frames=1, device_budget_bytes=0, log=lambda *a: None)
    np.testing.assert_array_equal(out2["map"], z["map"])


def test_sequence_map_full_size_vs_oracle():
    from pcl_augmentation_b200.semantic_segmentation.rich_map import drivable_area_map as srm
    labels = synth.load_config("ss")["insertion"]["placement_labels"]
    frames = _chained_frames([61, 62, 63, 64, 65, 66], synth.KITTI_SHAPE)
    frames += _chained_frames([67], synth.OS128_SHAPE) + _chained_frames([68], synth.SMALL_SHAPE)      # ragged
    frames.append((np.array([[1.0, 2.0, -1.7, 0.3]], dtype=np.float32), np.array([1], dtype=np.uint32), frames[0][2]))  # no surface point
    want = rmo.rich_map_ss(_oracle_frames(frames), labels)
    for chunk in (256, 4, 1):
        got = srm.sequence_map(frames, labels, chunk_frames=chunk)
        assert got["map"].dtype == np.float64
        np.testing.assert_array_equal(got["map"], want["map"])
        np.testing.assert_array_equal(got["move"], want["move"])
    assert {1.0, 2.0, 3.0} <= set(np.unique(want["map"]))


def test_sequence_map_projective_pose_and_label_precedence():
    """ss/rm:137 divides by the homogeneous coordinate; ss/rm:192-200 tests class 1 first, then 3, else 2."""
    from pcl_augmentation_b200.semantic_segmentation.rich_map import drivable_area_map as srm
    frames = _chained_frames([71, 72], synth.SMALL_SHAPE)
    T = frames[1][2].copy()
    T[3] = [0.0, 0.0, 1e-3, 2.0]
    frames[1] = (frames[1][0], frames[1][1], T)
    labels = {1: [40, 44], 2: [48, 70], 3: [44, 48]}               # 44 -> 1 (class 1 wins), 48 -> 3, 70 -> 2
    want = rmo.rich_map_ss(_oracle_frames(frames), labels)
    got = srm.sequence_map(frames, labels)
    np.testing.assert_array_equal(got["map"], want["map"])
    np.testing.assert_array_equal(got["move"], want["move"])
