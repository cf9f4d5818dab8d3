"""GPU tier: the stream-level primitives of the placement / occlusion / insertion stages (r3d_place_candidates,
r3d_obb_collide, r3d_occlude_mask, r3d_compact_insert — device pointers, caller-allocated, re-entrant per stream)
against the numpy oracle on the reference's own row layouts."""
import numpy as np
import pytest

from oracle import real3d_oracle as orc
from pcl_augmentation_b200 import boxes as bx
from pcl_augmentation_b200 import ops, synth
from tests.helpers import GOLDEN_SHAPE

pytestmark = pytest.mark.gpu


def scene_rows9(case):
    pcl5 = case.pcl5.copy()
    if case.task == "od":
        pcl5[pcl5[:, 4] != case.config["labels"]["Road"], 4] = 1            # od/ins:353-355
    pc = orc.add_space_for_spherical(pcl5)
    pc, mx, mn = orc.fill_spherical(pc)
    train, label, pc = orc.geometrical_front_view(pc, 112, 1440, mx, mn)
    return pc, train, label, mx, mn


@pytest.mark.parametrize("task", ["od", "ss"])
def test_place_candidates_matches_the_oracle_walk(task):
    """on-map flag and road level of every yaw candidate, incl. the carried z shift of the semseg walk"""
    case = synth.make_case(task, 1201, shape=GOLDEN_SHAPE, counts=None, n_cars=5, obj_range=(4.0, 16.0))
    classes = case.config["insertion"]["classes"]
    read = orc.read_label_line_ss if task == "ss" else orc.read_label_line_od
    for ci in (0, len(classes) - 1):
        cls = classes[ci]
        name, sample = case.db[cls][3]
        anno = read(str(sample["anno"]))
        pcl = np.array(sample["pcl"], dtype=np.float64)
        if task == "od":
            md = case.maps[case.config["insertion"]["placement"][cls]]
            surface = [case.config["labels"]["Road"]]
            flags, level = ops.place_candidates(pcl, bx.read_label_line_od(str(sample["anno"])), 360, "od",
                                                np.asarray(md["map"]).astype(np.uint8), (int(md["min_x"]), int(md["min_y"])),
                                                case.pcl5, surface)
            ground = orc._ground_rows_od(case.pcl5, surface[0])
        else:
            ok_map = case.config["insertion"]["placement"][cls]
            surface = [v for m in ok_map for v in case.config["insertion"]["placement_labels"][m]]
            mv = np.asarray(case.map_data["move"]).reshape(-1)
            flags, level = ops.place_candidates(pcl, bx.read_label_line_ss(str(sample["anno"])), 360, "ss",
                                                np.asarray(case.map_data["map"]).astype(np.uint8), (int(mv[0]), int(mv[1])),
                                                case.pcl5, surface, pose=case.pose, ok_map_values=ok_map)
            ground = orc._ground_rows_ss(case.pcl5, surface)
        # the oracle's ordered walk (closed-form rotation, carried shift for semseg)
        cos_k, sin_k = orc.yaw_tables(360)
        dz, n_on = 0.0, 0
        for k in range(1, 361):
            p = pcl.copy()
            p[:, 0] = cos_k[k] * pcl[:, 0] - sin_k[k] * pcl[:, 1]
            p[:, 1] = sin_k[k] * pcl[:, 0] + cos_k[k] * pcl[:, 1]
            p[:, 2] = pcl[:, 2] + dz
            if task == "od":
                on = orc.on_map_od(p, np.asarray(md["map"]), (int(md["min_x"]), int(md["min_y"])))
            else:
                on = orc.on_map_ss(p, np.asarray(case.map_data["map"]), np.asarray(case.map_data["move"]), case.pose, ok_map)
            assert bool(flags[k] & 1) == on, (cls, k)
            if not on:
                continue
            n_on += 1
            cx = cos_k[k] * anno["center"]["x"] - sin_k[k] * anno["center"]["y"]
            cy = sin_k[k] * anno["center"]["x"] + cos_k[k] * anno["center"]["y"]
            lv, ok = orc.road_level(ground, cx, cy)
            assert bool(flags[k] & 2) == ok, (cls, k)
            if ok:
                assert abs(level[k] - lv) <= 1e-9
                dz = lv - anno["center"]["z"]
        assert n_on > 0


@pytest.mark.parametrize("task", ["od", "ss"])
def test_obb_collide_batch_of_candidates_vs_oracle(task):
    case = synth.make_case(task, 1211, shape=GOLDEN_SHAPE, counts=None, n_cars=9, obj_range=(4.0, 16.0))
    scene, *_ = scene_rows9(case)
    read = orc.read_label_line_ss if task == "ss" else orc.read_label_line_od
    annos = [read(l) for l in case.box_lines]
    cls = case.config["insertion"]["classes"][-1]
    sample = case.db[cls][5][1]
    anno0 = read(str(sample["anno"]))
    pcl = np.array(sample["pcl"], dtype=np.float64)
    if task == "ss":
        ok_surface = [v for m in case.config["insertion"]["placement"][cls] for v in case.config["insertion"]["placement_labels"][m]]
    cands, want = [], []
    cos_k, sin_k = orc.yaw_tables(360)
    for k in range(1, 361, 7):
        p, a = orc.rotate_bounding_box(pcl.copy(), anno0, k, ss=task == "ss")
        dz = 0.02 * (k % 5)
        p[:, 2] += dz
        a = dict(a, center=dict(a["center"], z=a["center"]["z"] + dz))
        cands.append((cos_k[k], sin_k[k], dz, a))
        want.append(orc.collide_od(scene, annos, p, a) if task == "od" else orc.collide_ss(scene, annos, p, a, ok_surface))
    got = ops.obb_collide(scene, annos, pcl, cands, mode=task, pedestrian=(task == "od" and cls == "Pedestrian"),
                          ok_surface=ok_surface if task == "ss" else ())
    assert got.tolist() == want and 0 < sum(want) < len(want)


def test_occlude_mask_and_compact_insert_vs_oracle():
    case = synth.make_case("od", 1221, shape=GOLDEN_SHAPE, counts=[1, 1], n_cars=4, obj_range=(4.0, 16.0))
    scene, train, label, mx, mn = scene_rows9(case)
    s_train, _ = orc.smooth_out(train, label)
    done = 0
    for cls in case.config["insertion"]["classes"]:
        for j in range(6):
            sample = np.array(case.db[cls][j][1]["pcl"], dtype=np.float64)
            sample[:, 4] = 1
            keep, visible, vis = orc.occlude(scene, s_train, sample, 112, 1440, mx, mn)
            obj = orc.add_space_for_spherical(sample)
            obj, _, _ = orc.fill_spherical(obj)
            o_train, o_label, obj = orc.geometrical_front_view(obj, 112, 1440, mx, mn, sample=True)
            o_train, _ = orc.smooth_out(o_train, o_label)
            rows, n_vis, vis_gpu = ops.occlude_and_insert(scene, s_train, obj, o_train)
            np.testing.assert_array_equal(vis_gpu, vis)
            assert n_vis == len(visible)
            np.testing.assert_array_equal(rows[:len(rows) - n_vis], scene[keep])          # kept scene rows, in order
            np.testing.assert_array_equal(rows[len(rows) - n_vis:], visible)             # (pix_id, index) order
            done += n_vis > 0
    assert done >= 3
