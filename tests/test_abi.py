"""CPU tier: the C-ABI library loads and exports every symbol include/real3d_b200.h declares (no compute calls)."""
import os
import re

import pytest

from pcl_augmentation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "real3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(r3d_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert n in _lib.PROTOTYPES, f"{n} declared in the header but not bound in _lib.PROTOTYPES"
        assert getattr(lib, n) is not None
    for n in _lib.PROTOTYPES:
        assert n in names, f"{n} bound in _lib.py but not declared in include/real3d_b200.h"


def test_version_and_launch_counter():
    lib = _lib.load()
    assert lib.r3d_version() >= 100
    assert lib.r3d_launch_count() >= 0


def test_struct_layouts_match_header_constants():
    assert _lib.R3D_MAX_CLASSES == 16 and _lib.R3D_MAX_SURFACE == 8 and _lib.R3D_NUM_RADII == 50
    import ctypes as C
    assert C.sizeof(_lib.ClassCfg) == 4 * 5 + 4 * 8
    # EngineCfg: 13 int32 + 8 int32 + 1 int32 (+pad) + 50 doubles + 50 int32 + 16 class cfgs
    assert C.sizeof(_lib.EngineCfg) % 8 == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pcl_augmentation_b200 import ops
    import numpy as np
    with pytest.raises(_lib.Real3DError):
        ops.fill_spherical(np.zeros((4, 9)))
