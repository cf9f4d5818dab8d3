"""TEST INFRASTRUCTURE ONLY — compatibility shim that imports the UNMODIFIED reference from /root/reference.

Only ``oracle/make_golden.py`` and the ``tests/`` that pin the oracle against the real reference use this; it is
never imported by the product (``pcl_augmentation_b200``), ``bench.py`` or anything that runs on the GPU box
(``/root/reference`` does not exist there — every caller skips when it is absent).

What is patched (all results-neutral):
  * ``np.int``                       removed in numpy>=1.24 -> alias of ``int``
                                     (used at semantic_segmentation/Real3DAug/tools/find_spot.py:238, insertion.py:218)
  * ``Rotation.as_dcm/from_dcm``     removed in scipy>=1.6 -> ``as_matrix/from_matrix``
                                     (tools/cut_bbox.py:28, tools/find_spot.py:83,90,213)
  * ``skimage`` (not installed)      stub modules: ``img_as_ubyte``, ``rectangle``, ``disk``, ``closing``, ``dilation`` built on
                                     ``scipy.ndimage.grey_dilation/grey_erosion`` (mode='reflect', which is what
                                     ``skimage.morphology.closing`` wraps).  scikit-image is an un-pinned third-party
                                     dependency of the reference (tools/closing.py:2-4): parity of the closing step is
                                     pinned by this definition (binary 5x3 max-then-min, out-of-image ignored).
  * semseg ``insertion.py`` uses the undefined names ``ROAD_INDEXES`` / ``MAC`` (insertion.py:209,247) ->
    injected as ``[40, 44, 48]`` (the value in object_detection/Real3DAug/tools/find_spot.py:14) / ``False``.
  * ``glob.glob`` -> sorted, ``random.shuffle`` / ``input`` -> scripted (only inside ``run_main``).
"""
from __future__ import annotations

import builtins
import contextlib
import glob as _glob
import importlib
import io
import os
import random as _random
import runpy
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("R3D_REFERENCE_ROOT", "/root/reference")
TREES = {"od": "object_detection/Real3DAug", "ss": "semantic_segmentation/Real3DAug"}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, TREES["od"], "insertion.py"))


def _install_compat():
    if not hasattr(np, "int"):
        np.int = int                                                     # noqa: NPY001
    from scipy.spatial.transform import Rotation
    if not hasattr(Rotation, "as_dcm"):
        Rotation.as_dcm = Rotation.as_matrix
        Rotation.from_dcm = Rotation.from_matrix
    try:
        import skimage  # noqa: F401
    except ImportError:
        from scipy import ndimage

        sk = types.ModuleType("skimage")
        util = types.ModuleType("skimage.util")
        morph = types.ModuleType("skimage.morphology")
        sio = types.ModuleType("skimage.io")

        def img_as_ubyte(a):
            a = np.asarray(a)
            if a.dtype == np.uint8:
                return a
            return np.round(a * 255.0).astype(np.uint8)

        def rectangle(nrows, ncols, dtype=np.uint8):
            return np.ones((nrows, ncols), dtype=dtype)

        def closing(image, footprint=None, out=None):
            fp = np.asarray(footprint, dtype=bool)
            dil = ndimage.grey_dilation(image, footprint=fp)
            return ndimage.grey_erosion(dil, footprint=fp)

        def disk(radius, dtype=np.uint8):
            L = np.arange(-radius, radius + 1)
            X, Y = np.meshgrid(L, L)
            return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)

        def dilation(image, footprint=None, out=None):
            return ndimage.grey_dilation(image, footprint=np.asarray(footprint, dtype=bool))

        def imread(fname, *a, **k):
            from PIL import Image
            return np.asarray(Image.open(fname))

        sio.imread = imread
        util.img_as_ubyte = img_as_ubyte
        morph.rectangle = rectangle
        morph.closing = closing
        morph.disk = disk
        morph.dilation = dilation
        sk.util, sk.morphology, sk.io = util, morph, sio
        sys.modules.update({"skimage": sk, "skimage.util": util, "skimage.morphology": morph, "skimage.io": sio})


def _purge():
    for k in list(sys.modules):
        if k == "tools" or k.startswith("tools.") or k == "insertion":
            del sys.modules[k]


@contextlib.contextmanager
def _tree_on_path(which):
    path = os.path.join(REFERENCE_ROOT, TREES[which])
    _purge()
    sys.path.insert(0, path)
    try:
        yield path
    finally:
        sys.path.remove(path)
        _purge()


class RefModules:
    """Handles to one tree's modules (``insertion``, ``find_spot``, ``closing``, ``cut_bbox``, ``datasets``)."""

    def __init__(self, which):
        _install_compat()
        with _tree_on_path(which):
            self.insertion = importlib.import_module("insertion")
            self.find_spot = importlib.import_module("tools.find_spot")
            self.closing = importlib.import_module("tools.closing")
            self.cut_bbox = importlib.import_module("tools.cut_bbox")
            self.datasets = importlib.import_module("tools.datasets")
        if which == "ss":
            self.insertion.ROAD_INDEXES = [40, 44, 48]
            self.insertion.MAC = False
        self.which = which

    def set_image_size(self, rows, cols):
        self.insertion.NUMROW = rows
        self.insertion.NUMCOLUMN = cols


_CACHE = {}


def load(which) -> RefModules:
    if which not in _CACHE:
        _CACHE[which] = RefModules(which)
    return _CACHE[which]


class FreshDict(dict):
    """Stand-in for ``np.lib.npyio.NpzFile``: every ``[...]`` hands out a fresh copy, as an NpzFile re-reads the array
    from disk on each access (so the reference's in-place edits never persist between calls)."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        return v.copy() if isinstance(v, np.ndarray) else v


def run_main(which, cwd, inputs=(), shuffle_fn=None, quiet=True):
    """Execute the reference ``insertion.py`` as ``__main__`` (unmodified) with ``cwd`` as working directory."""
    _install_compat()
    it = iter(inputs)
    old_input, old_glob, old_shuffle, old_cwd = builtins.input, _glob.glob, _random.shuffle, os.getcwd()
    builtins.input = lambda *a: next(it)
    _glob.glob = lambda *a, **k: sorted(old_glob(*a, **k))
    if shuffle_fn is not None:
        _random.shuffle = shuffle_fn
    init = {"ROAD_INDEXES": [40, 44, 48], "MAC": False} if which == "ss" else {}
    out = io.StringIO()
    try:
        os.chdir(cwd)
        with _tree_on_path(which) as path:
            with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
                runpy.run_path(os.path.join(path, "insertion.py"), init_globals=init, run_name="__main__")
    finally:
        os.chdir(old_cwd)
        builtins.input, _glob.glob, _random.shuffle = old_input, old_glob, old_shuffle
    return out.getvalue()


def run_rich_map_od(cwd, quiet=True):
    """Execute the reference's ``object_detection/rich_map/single_drivable_area_map.py`` as ``__main__`` (unmodified);
    it reads ``../config/KITTI.yaml`` relative to ``cwd`` and imports ``object_detection.Real3DAug.tools.datasets``."""
    _install_compat()
    old_cwd = os.getcwd()
    out = io.StringIO()
    for k in [k for k in sys.modules if k == "object_detection" or k.startswith("object_detection.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        os.chdir(cwd)
        with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
            runpy.run_path(os.path.join(REFERENCE_ROOT, "object_detection/rich_map/single_drivable_area_map.py"),
                           run_name="__main__")
    finally:
        os.chdir(old_cwd)
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "object_detection" or k.startswith("object_detection.")]:
            del sys.modules[k]
    return out.getvalue()


def run_rich_map_ss(cwd, inputs, quiet=True):
    """Execute the reference's ``semantic_segmentation/rich_map/drivable_area_map.py`` as ``__main__`` (unmodified);
    it reads ``../config/semantic-kitti.yaml`` relative to ``cwd``, imports ``semantic_segmentation.Real3DAug.tools.datasets``
    and asks for the dataset, the sequence and the frame order on ``input()`` (answers: ``inputs``)."""
    _install_compat()
    old_cwd, old_input, old_glob = os.getcwd(), builtins.input, _glob.glob
    it = iter(inputs)
    builtins.input = lambda *a: next(it)
    _glob.glob = lambda *a, **k: sorted(old_glob(*a, **k))
    out = io.StringIO()
    mods = lambda: [k for k in sys.modules if k == "semantic_segmentation" or k.startswith("semantic_segmentation.")]
    for k in mods():
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        os.chdir(cwd)
        with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
            runpy.run_path(os.path.join(REFERENCE_ROOT, "semantic_segmentation/rich_map/drivable_area_map.py"),
                           run_name="__main__")
    finally:
        os.chdir(old_cwd)
        builtins.input, _glob.glob = old_input, old_glob
        sys.path.remove(REFERENCE_ROOT)
        for k in mods():
            del sys.modules[k]
    return out.getvalue()


def _run_script(rel_path, extra_paths, cwd, inputs, prepare=None, quiet=True):
    """Execute a reference script as ``__main__`` (unmodified) from ``cwd`` with scripted ``input()`` answers,
    ``REFERENCE_ROOT`` + ``extra_paths`` (relative to it) on ``sys.path``."""
    _install_compat()
    old_cwd, old_input, old_glob = os.getcwd(), builtins.input, _glob.glob
    it = iter(inputs)
    builtins.input = lambda *a: next(it)
    _glob.glob = lambda *a, **k: sorted(old_glob(*a, **k))
    out = io.StringIO()
    roots = ("semantic_segmentation", "object_detection", "tools", "cut_bbox", "cutout")
    mods = lambda: [k for k in sys.modules if k.split(".")[0] in roots]
    for k in mods():
        del sys.modules[k]
    paths = [REFERENCE_ROOT] + [os.path.join(REFERENCE_ROOT, p) for p in extra_paths]
    for p in reversed(paths):
        sys.path.insert(0, p)
    try:
        os.chdir(cwd)
        if prepare is not None:
            prepare()
        with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
            runpy.run_path(os.path.join(REFERENCE_ROOT, rel_path), run_name="__main__")
    finally:
        os.chdir(old_cwd)
        builtins.input, _glob.glob = old_input, old_glob
        for p in paths:
            sys.path.remove(p)
        for k in mods():
            del sys.modules[k]
    return out.getvalue()


def run_cut_objects_od(cwd, quiet=True):
    """``object_detection/cut_object/object_cut_out.py`` (unmodified): reads ``../config/KITTI.yaml``, imports its
    sibling modules ``cut_bbox`` / ``cutout`` and ``object_detection.Real3DAug.tools.datasets``."""
    return _run_script("object_detection/cut_object/object_cut_out.py", ["object_detection/cut_object"], cwd, [], quiet=quiet)


def _accept_subdirectories_keyword():
    """Reference defect, results-neutral patch: ``cut_out.py:98`` calls ``delete_item(0, subdirectoties=False)``, a
    keyword only ``Waymo.delete_item`` accepts (ss/ds:266) — with ``SemanticKITTI`` the script raises TypeError."""
    ds = importlib.import_module("semantic_segmentation.Real3DAug.tools.datasets")
    plain = ds.SemanticKITTI.delete_item
    ds.SemanticKITTI.delete_item = lambda self, idx, subdirectoties=True: plain(self, idx)


def run_cut_objects_ss(cwd, inputs, quiet=True):
    """``semantic_segmentation/cut_object/cut_out.py`` (unmodified but for the keyword patch above).  Second reference
    defect: ``cut_out.py:55`` / ``filter_objects.py:56`` format the ``input()`` string with ``:02d`` (ValueError for any
    str), so the scripted answer to the sequence prompt has to be an ``int`` (``inputs = ["1", 0, "no"]``)."""
    return _run_script("semantic_segmentation/cut_object/cut_out.py", ["semantic_segmentation/cut_object"], cwd, inputs,
                       prepare=_accept_subdirectories_keyword, quiet=quiet)


def run_filter_objects_ss(cwd, inputs, quiet=True):
    """``semantic_segmentation/cut_object/filter_objects.py`` (unmodified)."""
    return _run_script("semantic_segmentation/cut_object/filter_objects.py", ["semantic_segmentation/cut_object"], cwd,
                       inputs, prepare=_accept_subdirectories_keyword, quiet=quiet)
