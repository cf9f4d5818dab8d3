"""CPU tier, SURVEY §8f row 4: the cut-object oracle against what the UNMODIFIED reference scripts
(object_cut_out.py, cut_out.py, filter_objects.py) wrote for the seeded synthetic datasets (tests/golden)."""
import json

import numpy as np

from oracle import cut_objects_oracle as coo
from pcl_augmentation_b200 import synth
from tests.helpers import (KITTI_CALIB_LINES, KITTI_IMAGE_SHAPE, cut_object_cases, load_golden)


def golden_samples(g, key):
    names, annos = json.loads(str(g[key + "_names"])), json.loads(str(g[key + "_annos"]))
    off = np.concatenate(([0], np.cumsum(g[key + "_counts"])))
    return {n: (a, g[key + "_pcl"][off[i]:off[i + 1]]) for i, (n, a) in enumerate(zip(names, annos))}


def calib_from_lines(tmp_path):
    p = tmp_path / "calib.txt"
    p.write_text("\n".join(KITTI_CALIB_LINES) + "\n")
    return coo.read_calib(str(p))


def frame_points(case):
    return np.hstack((case.pcl5[:, :4].astype(np.float32), case.pcl5[:, 4:5]))        # KITTI.__getitem__ (od/ds:62-66)


def test_cut_objects_od_matches_reference_script(tmp_path):
    g = load_golden("cut_objects_od")
    cases = cut_object_cases("od")
    calib = calib_from_lines(tmp_path)
    got = {}
    for i, case in enumerate(cases):
        assert synth.array_digest(case.pcl5) == str(g[f"digest{i}"])
        for cls, name, anno, pcl in coo.cut_objects_od(frame_points(case), [l + "\n" for l in case.box_lines], calib,
                                                       KITTI_IMAGE_SHAPE, case.config, f"{i:06d}"):
            got.setdefault(cls, {})[name] = (anno, pcl)
    n = 0
    for cls in cases[0].config["insertion"]["classes"]:
        want = golden_samples(g, cls)
        assert sorted(got.get(cls, {})) == sorted(want)
        for name, (anno, pcl) in want.items():
            assert got[cls][name][0] == anno
            assert got[cls][name][1].dtype == np.float64
            np.testing.assert_array_equal(got[cls][name][1], pcl)
            n += 1
    assert n >= 10


def test_cut_objects_ss_and_filter_match_reference_scripts():
    g = load_golden("cut_objects_ss")
    cases = cut_object_cases("ss")
    cfg = cases[0].config
    got = {}
    for i, case in enumerate(cases):
        assert synth.array_digest(case.pcl5) == str(g[f"digest{i}"])
        for folder, name, anno, pcl in coo.cut_objects_ss(frame_points(case), [l + "\n" for l in case.box_lines], cfg, "00",
                                                          f"{i:06d}"):
            got.setdefault(folder, {})[name] = (anno, pcl)
    removed = 0
    for cls in cfg["insertion"]["classes"]:
        folder = cfg["labels"][cls]
        want = golden_samples(g, folder)
        assert sorted(got.get(folder, {})) == sorted(want)
        for name, (anno, pcl) in want.items():
            assert got[folder][name][0] == anno
            np.testing.assert_array_equal(got[folder][name][1], pcl)
        # filter_objects.py: per distance bucket (the '_<ddd>_m' suffix of the file name)
        kept = set(want)
        for d in range(100):
            bucket = [(n, a, len(p)) for n, (a, p) in want.items() if n.endswith(f"_{d:03d}_m")]
            if bucket:
                kept -= set(coo.filter_objects(bucket))
        assert sorted(kept) == json.loads(str(g[folder + "_kept"]))
        removed += len(want) - len(kept)
    assert removed >= 10
