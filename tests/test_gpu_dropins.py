"""GPU tier: the function-level drop-ins (same names / arguments as the reference's tools/find_spot.py, insertion.py)
against the golden fixtures produced by the unmodified reference and against the oracle."""
import copy

import numpy as np
import pytest

from oracle import real3d_oracle as orc
from pcl_augmentation_b200.object_detection.Real3DAug import insertion as od_ins
from pcl_augmentation_b200.object_detection.Real3DAug.tools import find_spot as od_fs
from pcl_augmentation_b200.semantic_segmentation.Real3DAug import insertion as ss_ins
from pcl_augmentation_b200.semantic_segmentation.Real3DAug.tools import find_spot as ss_fs
from tests.helpers import case_from_golden, load_golden

pytestmark = pytest.mark.gpu


def _scene9(ins, case):
    pcl5 = case.pcl5.copy()
    if case.task == "od":
        pcl5[pcl5[:, 4] != case.config["labels"]["Road"], 4] = 1
    pc = ins.add_space_for_spherical(pcl5)
    pc, mx, mn = ins.fill_spherical(pc)
    _, _, pc = ins.geometrical_front_view(pc, 112, 1440, mx, mn)
    return pcl5, pc


@pytest.mark.parametrize("task", ["od", "ss"])
def test_find_possible_places_dropin_vs_reference_golden(task):
    g = load_golden(f"fn_places_{task}")
    spec, case = case_from_golden(g)
    ins, fs = (od_ins, od_fs) if task == "od" else (ss_ins, ss_fs)
    original, scene = _scene9(ins, case)
    annos = [fs.read_label_line(l) for l in case.box_lines]
    classes = case.config["insertion"]["classes"]
    if task == "ss":
        md = {"map": case.map_data["map"].copy(), "move": case.map_data["move"]}
        map_arr, map_move = ss_ins.addjust_map_2(md, scene, case.pose)
        np.testing.assert_array_equal(np.argwhere(map_arr == 4).astype(np.int32), g["map_adjusted_cells"])
    total = 0
    for ci, cls in enumerate(classes):
        for j in range(2):
            name, sample = case.db[cls][j]
            sd = {"pcl": sample["pcl"].copy(), "anno": sample["anno"]}
            if task == "od":
                placement = case.config["insertion"]["placement"][cls]
                pcls, ans, rots = fs.find_possible_places(scene, annos, sd, case.maps[placement], original, case.config)
            else:
                pcls, ans, rots = fs.find_possible_places(scene, annos, sd, map_arr, map_move, original, case.pose,
                                                          case.config)
            key = f"c{ci}_s{j}"
            np.testing.assert_array_equal(np.array(rots, dtype=np.int32), g[key + "_rots"])
            total += len(rots)
            if rots:
                for n, p in enumerate(g[key + "_pick"]):
                    np.testing.assert_allclose(pcls[p][:, :3], g[key + "_xyz"][n], rtol=0, atol=1e-9)
                    np.testing.assert_array_equal(pcls[p][:, 3:], sd["pcl"][:, 3:])
                    box = g[key + "_box"][n]
                    c = ans[p]["center"]
                    np.testing.assert_allclose([c["x"], c["y"], c["z"]], box[:3], rtol=0, atol=1e-9)
                    np.testing.assert_allclose(orc.box_matrix(ans[p]), orc.R.from_quat(box[3:7]).as_matrix(), rtol=0, atol=1e-9)
    assert total > 50


def test_rotate_correct_height_check_bounding_box_vs_oracle():
    g = load_golden("fn_places_od")
    spec, case = case_from_golden(g)
    original, scene = _scene9(od_ins, case)
    annos = [od_fs.read_label_line(l) for l in case.box_lines]
    name, sample = case.db["Cyclist"][1]
    pcl_a, pcl_b = sample["pcl"].copy(), sample["pcl"].copy()
    anno_a = od_fs.read_label_line(str(sample["anno"]))
    anno_b = orc.read_label_line_od(str(sample["anno"]))
    for rot in (1, 37.5, 180):
        pcl_a, anno_a = od_fs.rotate_bounding_box(pcl_a, anno_a, rot)
        pcl_b, anno_b = orc.rotate_bounding_box(pcl_b, anno_b, rot)
        np.testing.assert_allclose(pcl_a, pcl_b, rtol=0, atol=1e-12)
        np.testing.assert_allclose(list(anno_a["center"].values()), list(anno_b["center"].values()), rtol=0, atol=1e-12)
    # road level around a few centres on and off the road
    ground = orc._ground_rows_od(original, 40)
    for cx, cy in [(10.0, 0.5), (-20.0, -2.0), (3.0, 30.0), (200.0, 200.0)]:
        a = copy.deepcopy(anno_a)
        a["center"]["x"], a["center"]["y"] = cx, cy
        p = pcl_a.copy()
        p2, a2, ok = od_fs.correct_height(original, p, a, case.config)
        level, ok_ref = orc.road_level(ground, cx, cy)
        assert ok == ok_ref
        if ok:
            assert a2["center"]["z"] == level
            np.testing.assert_allclose(p2[:, 2], pcl_a[:, 2] + (level - anno_a["center"]["z"]), rtol=0, atol=1e-12)
    # collision predicate on a few candidates
    pcls, ans, rots = orc.find_possible_places_od(scene, annos, orc.FreshDict(sample), case.maps["Road"], original,
                                                  case.config, mode="closed")
    for p, a in list(zip(pcls, ans))[:3]:
        assert od_fs.check_bounding_box(scene, annos, p, a) is True
        shifted = copy.deepcopy(a)
        shifted["center"]["x"], shifted["center"]["y"] = annos[0]["center"]["x"], annos[0]["center"]["y"]
        moved = p.copy()
        moved[:, 0] += shifted["center"]["x"] - a["center"]["x"]
        moved[:, 1] += shifted["center"]["y"] - a["center"]["y"]
        assert od_fs.check_bounding_box(scene, annos, moved, shifted) == (not orc.collide_od(scene, annos, moved, shifted))
