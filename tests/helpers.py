"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

from pcl_augmentation_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_SHAPE = synth.ScanShape(32, 600, 2.0, -24.8)
CASE_DEFAULTS = dict(shape=GOLDEN_SHAPE, n_per_class=100, obj_range=(4.0, 16.0))


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def case_from_spec(spec):
    kw = dict(CASE_DEFAULTS)
    kw.update({k: v for k, v in spec.items() if k not in ("task", "seed")})
    return synth.make_case(spec["task"], spec["seed"], **kw)


def case_from_golden(g):
    spec = json.loads(str(g["meta"]))["spec"]
    case = case_from_spec(spec)
    assert synth.case_digest(case) == str(g["digest"]), "regenerated synthetic inputs differ from the golden run"
    return spec, case


def parse_inserted(text):
    """'<name> with rotation: <rot>' lines of added_objects/<frame>.txt (od/ins:537)."""
    out = []
    for line in text.splitlines():
        if line.strip():
            name, rot = line.split(" with rotation: ")
            out.append((name, int(rot)))
    return out


RICH_MAP_SS_SEEDS = (41, 42, 43, 44)


def rich_map_ss_cases():
    """Four frames of one synthetic sequence: the poses are chained so that the frames overlap (cells written by
    several frames and by several surface classes: the last-writer and the sticky-sidewalk rules both matter)."""
    cases = [case_from_spec(dict(task="ss", seed=s, counts=[1] * 8, n_cars=c)) for s, c in zip(RICH_MAP_SS_SEEDS, (0, 3, 5, 2))]
    base = cases[0].pose
    for i, case in enumerate(cases[1:], 1):
        a = 0.35 * i
        step = np.eye(4)
        step[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        step[:3, 3] = [7.3 * i, -2.1 * i, 0.05 * i]
        case.pose = base @ step
    return cases
