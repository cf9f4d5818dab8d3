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


# ----------------------------------------------------------------------- stream-level placement / occlusion / insertion
def obb_collide(scene_pcl, scene_annos, sample_pcl, candidates, mode='od', pedestrian=False, ok_surface=()):
    """check_bounding_box (od/fs:109-135, ss/fs:79-104) for several candidate placements of one object in ONE call.

    ``candidates``: list of ``(cos, sin, dz, annotation)`` — candidate k's points are ``sample_pcl`` turned by
    (cos, sin) about the sensor z axis and lifted by dz, its box is ``annotation``.  Returns a bool array: True = the
    candidate collides (with an obstacle scene point inside its box, or with a scene box around one of its points)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    k = len(candidates)
    if k == 0:
        return np.zeros(0, dtype=bool)
    scene = _dev(scene_pcl, np.float64)
    obj = _dev(sample_pcl, np.float64)
    sboxes = _dev(np.array([box_record(a) for a in scene_annos], dtype=np.float64).reshape(-1, 16), np.float64)
    cand4 = _dev(np.array([[c, s, dz, 0.0] for c, s, dz, _ in candidates], dtype=np.float64), np.float64)
    cboxes = _dev(np.array([box_record(a) for _, _, _, a in candidates], dtype=np.float64), np.float64)
    ok = np.asarray(list(ok_surface), dtype=np.int32)
    out = torch.empty(k, dtype=torch.uint8, device="cuda")
    _lib.check(lib.r3d_obb_collide(scene.data_ptr(), len(scene_pcl), sboxes.data_ptr() if len(scene_annos) else None,
                                   len(scene_annos), obj.data_ptr(), len(sample_pcl), sample_pcl.shape[1], cand4.data_ptr(),
                                   cboxes.data_ptr(), k, 0 if mode == 'od' else 1, 1 if pedestrian else 0,
                                   ok.ctypes.data if len(ok) else None, len(ok), out.data_ptr(), _stream()), "obb_collide")
    return out.cpu().numpy().astype(bool)


def occlude_and_insert(scene_pcl, scene_train, sample_pcl9, sample_train):
    """The occlusion step of insertion.py (od/ins:486-501, 545) on already projected and smoothed inputs:
    ``scene_pcl`` / ``sample_pcl9`` are N x 9 / M x 9 working rows with their pix ids in column 8, ``scene_train`` /
    ``sample_train`` the smoothed range images.  Returns (new scene rows = kept scene rows + visible object rows in
    (pix_id, index) order, number of visible object rows, vis_px bool image)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    h, w = scene_train.shape
    n, m = len(scene_pcl), len(sample_pcl9)
    scene = _dev(scene_pcl, np.float64)
    obj = _dev(sample_pcl9, np.float64)
    st, ot = _dev(scene_train, np.float64), _dev(sample_train, np.float64)
    skeep = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
    okeep = torch.empty(max(m, 1), dtype=torch.uint8, device="cuda")
    vis = torch.empty(h * w, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(2, dtype=torch.int32, device="cuda")
    _lib.check(lib.r3d_occlude_mask(scene.data_ptr(), n, obj.data_ptr(), m, st.data_ptr(), ot.data_ptr(), h * w,
                                    skeep.data_ptr(), okeep.data_ptr(), vis.data_ptr(), counts.data_ptr(), _stream()),
               "occlude_mask")
    out = torch.empty((n + m, 9), dtype=torch.float64, device="cuda")
    n_out = torch.zeros(2, dtype=torch.int64, device="cuda")
    cap = 1
    while cap < max(m, 1):
        cap <<= 1
    scratch = torch.empty(cap, dtype=torch.int64, device="cuda")
    _lib.check(lib.r3d_compact_insert(scene.data_ptr(), skeep.data_ptr(), n, obj.data_ptr(), okeep.data_ptr(), m,
                                      out.data_ptr(), n_out.data_ptr(), scratch.data_ptr(), cap, _stream()), "compact_insert")
    rows, kept = (int(v) for v in n_out.cpu().numpy())
    return out[:rows].cpu().numpy(), rows - kept, vis.cpu().numpy().reshape(h, w).astype(bool)


def place_candidates(sample_pcl, sample_anno, yaw_steps, task, map_u8, map_move, original_pcl5, surface_labels, pose=None,
                     ok_map_values=()):
    """A5 + A6 + A7 of find_possible_places (od/fs:263-285, ss/fs:229-250) for all yaw candidates of one cut object.
    Returns (flags uint8 [yaw_steps + 1]: bit0 on the map, bit1 road level found; road level float64 [yaw_steps + 1])."""
    from . import boxes as bx
    torch = _lib.require_cuda()
    lib = _lib.load()
    obj = _dev(sample_pcl, np.float64)
    ground = _dev(np.ascontiguousarray(original_pcl5[:, :5]), np.float64)
    cos_k, sin_k = bx.yaw_tables(yaw_steps)
    ck, sk = _dev(cos_k, np.float64), _dev(sin_k, np.float64)
    m8 = _dev(np.ascontiguousarray(map_u8, dtype=np.uint8), np.uint8)
    box8 = np.ascontiguousarray(bx.object_box_record(sample_anno), dtype=np.float64)
    r2, ok = bx.search_radii()
    lab = np.asarray(list(surface_labels), dtype=np.int32)
    mask = 0
    for v in ok_map_values:
        mask |= 1 << int(v)
    flags = torch.zeros(yaw_steps + 1, dtype=torch.uint8, device="cuda")
    level = torch.zeros(yaw_steps + 1, dtype=torch.float64, device="cuda")
    p16 = np.ascontiguousarray(pose, dtype=np.float64) if pose is not None else None
    _lib.check(lib.r3d_place_candidates(obj.data_ptr(), len(sample_pcl), sample_pcl.shape[1], box8.ctypes.data, yaw_steps,
                                        ck.data_ptr(), sk.data_ptr(), 0 if task == 'od' else 1, m8.data_ptr(), map_u8.shape[0],
                                        map_u8.shape[1], int(map_move[0]), int(map_move[1]), p16.ctypes.data if p16 is not None else None,
                                        mask, ground.data_ptr(), len(original_pcl5), lab.ctypes.data, len(lab), r2.ctypes.data,
                                        ok.ctypes.data, flags.data_ptr(), level.data_ptr(), _stream()), "place_candidates")
    return flags.cpu().numpy(), level.cpu().numpy()
