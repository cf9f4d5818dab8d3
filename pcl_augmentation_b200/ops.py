"""Operator-level drop-ins of the reference's projection / closing / box-cut functions, computed in CUDA.

Same names, argument meaning, in-place mutation and return values as the reference
(``add_space_for_spherical`` / ``fill_spherical`` / ``geometrical_front_view`` od/ins:55-130, ``class_closing`` /
``smooth_out`` cl:9-62, ``cut_bounding_box`` cb:7-68): numpy arrays in, numpy arrays out.  PyTorch only provides the
device buffers and the current stream; the math is in ``libreal3d_b200.so`` (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .boxes import box_record

NUMROW = 112                # od/ins:21
NUMCOLUMN = 360 * 4         # od/ins:22


def _dev(arr, dtype):
    torch = _lib.require_cuda()
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).cuda()


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def add_space_for_spherical(point_cloud):
    """N x 5 -> N x 9 working rows, unset columns = -1 (od/ins:55-65).  Pure layout change, done on the host."""
    out = np.ones((len(point_cloud), 9)) * -1
    out[:, 0:3] = point_cloud[:, 0:3]
    out[:, 6:8] = point_cloud[:, 3:5]
    return out


def fill_spherical(point_cloud):
    """r / azimuth / elevation written in place to columns 3..5; returns (pc, max_el, min_el) (od/ins:68-82)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    assert point_cloud.ndim == 2 and point_cloud.shape[1] == 9 and point_cloud.dtype == np.float64
    d = _dev(point_cloud, np.float64)
    mm = torch.empty(2, dtype=torch.float64, device="cuda")
    _lib.check(lib.r3d_fill_spherical(d.data_ptr(), len(point_cloud), mm.data_ptr(), _stream()), "fill_spherical")
    host = d.cpu().numpy()
    point_cloud[:, 3:6] = host[:, 3:6]
    mx, mn = mm.cpu().numpy()
    if len(point_cloud) == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")   # np.min on empty
    return point_cloud, np.float64(mx), np.float64(mn)


def geometrical_front_view(point_cloud, num_row, num_column, max_elevation_angle, min_elevation_angle, sample=False):
    """Spherical z-buffer (od/ins:85-130): returns (train, label, point_cloud); pix_id goes to column 8."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    assert point_cloud.ndim == 2 and point_cloud.shape[1] == 9 and point_cloud.dtype == np.float64
    d = _dev(point_cloud, np.float64)
    px = num_row * num_column
    train = torch.empty(px, dtype=torch.float64, device="cuda")
    label = torch.empty(px, dtype=torch.float64, device="cuda")
    zbuf = torch.empty(px, dtype=torch.int64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.r3d_project_zbuffer(d.data_ptr(), len(point_cloud), num_row, num_column, NUMCOLUMN,
                                       float(max_elevation_angle), float(min_elevation_angle), 1 if sample else 0,
                                       train.data_ptr(), label.data_ptr(), zbuf.data_ptr(), status.data_ptr(),
                                       _stream()), "geometrical_front_view")
    if int(status.item()) != 0:
        raise AssertionError("Rows / columns in FoV went something wrong.")          # od/ins:111-113
    if len(point_cloud):
        point_cloud[:, 8] = d[:, 8].cpu().numpy()
    return (train.cpu().numpy().reshape(num_row, num_column), label.cpu().numpy().reshape(num_row, num_column),
            point_cloud)


def _close_fill(original_train, original_label, want_closed):
    torch = _lib.require_cuda()
    lib = _lib.load()
    h, w = original_label.shape
    t_in = _dev(original_train, np.float64)
    l_in = _dev(original_label, np.float64)
    t_out = torch.empty_like(t_in)
    l_out = torch.empty_like(l_in)
    closed = torch.empty((h, w), dtype=torch.uint8, device="cuda") if want_closed else None
    _lib.check(lib.r3d_close_fill(t_in.data_ptr(), l_in.data_ptr(), h, w, t_out.data_ptr(), l_out.data_ptr(),
                                  closed.data_ptr() if want_closed else None, _stream()), "smooth_out")
    return t_out, l_out, closed


def class_closing(original_label):
    """closing(rectangle(5, 3)) of the occupancy as a uint8 0/255 image (cl:9-23)."""
    _, _, closed = _close_fill(np.zeros_like(original_label, dtype=np.float64), original_label, True)
    return closed.cpu().numpy()


def smooth_out(original_train, original_label):
    """Closing + neighbour-mean hole fill of the range image (cl:26-62): returns (train, label)."""
    t, l, _ = _close_fill(original_train, original_label, False)
    return t.cpu().numpy(), l.cpu().numpy()


def cut_bounding_box_mask(point_cloud, annotation, annotation_move=(0, 0, 0)):
    torch = _lib.require_cuda()
    lib = _lib.load()
    n = len(point_cloud)
    if n == 0:
        return np.zeros(0, dtype=bool)
    d = _dev(point_cloud, np.float64)
    mask = torch.empty(n, dtype=torch.uint8, device="cuda")
    box = np.ascontiguousarray(box_record(annotation, annotation_move))
    _lib.check(lib.r3d_cut_bounding_box(d.data_ptr(), n, point_cloud.shape[1], box.ctypes.data, mask.data_ptr(),
                                        _stream()), "cut_bounding_box")
    return mask.cpu().numpy().astype(bool)


def cut_bounding_box(point_cloud, annotation, annotation_move=[0, 0, 0]):
    """Rows of ``point_cloud`` strictly inside the oriented box, z measured from the box bottom (cb:7-68)."""
    return point_cloud[cut_bounding_box_mask(point_cloud, annotation, annotation_move)]
