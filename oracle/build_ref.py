"""TEST / BASELINE INFRASTRUCTURE ONLY — recipe that makes the UNMODIFIED reference available to the GPU box.

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no build step): the files of the hot path are copied, byte for byte, from
/root/reference into ``oracle/_ref/`` with their relative paths.  ``oracle/_ref/`` is git-ignored (reference sources
never enter the history) but NOT gpurun-ignored, so it travels with the snapshot like a built ``.so``.  The only user is
``bench.py --impl reference`` (`cpu_baseline.kind = "reference"`): ``oracle/shim.py`` imports the tree through
``R3D_REFERENCE_ROOT`` exactly as ``oracle/make_golden.py`` does from /root/reference in this container.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
# the per-scan augmentation loop and what it imports
FILES = [f"{tree}/Real3DAug/{rel}" for tree in ("object_detection", "semantic_segmentation")
         for rel in ("__init__.py", "insertion.py", "tools/closing.py", "tools/cut_bbox.py", "tools/datasets.py",
                     "tools/find_spot.py")]


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/build_ref.py: {SRC} absent, keeping oracle/_ref as it is "
                  f"({'present' if available() else 'absent'})")
        return available()
    n = 0
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
            n += 1
    if verbose:
        print(f"oracle/build_ref.py: {len(FILES)} reference files under oracle/_ref ({n} copied)")
    return available()


def available():
    return os.path.isfile(os.path.join(DST, "object_detection", "Real3DAug", "insertion.py"))


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
