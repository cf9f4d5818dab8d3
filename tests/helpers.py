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
