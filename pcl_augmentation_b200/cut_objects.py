"""Batched point-in-box cutting for cut-object database building: every ground-truth box of every
frame in one pass over the frame's points (``csrc/r3d_cutdb.cu``).  The two script drop-ins
(``object_detection/cut_object/object_cut_out.py``, ``semantic_segmentation/cut_object/cut_out.py``) sit on top.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from . import boxes as bx

MAX_BOXES_PER_FRAME = 64
EMIT_ANY, EMIT_NONE = -1, -2


class BoxCut:
    """Result for one box: points strictly inside (``count_inside``), those the camera sees (``count_fov``, 0 without a
    camera), and the emitted points in their original order (``xyzi`` float32 M x 4, ``labels`` uint32 M)."""
    __slots__ = ("count_inside", "count_fov", "xyzi", "labels", "index")

    def __init__(self, count_inside, count_fov, xyzi, labels, index):
        self.count_inside, self.count_fov, self.xyzi, self.labels, self.index = count_inside, count_fov, xyzi, labels, index


def camera_record(calib, img_shape):
    """27 doubles for the C ABI from a KITTI calibration (float32 ``Tr_velo2cam`` 3 x 4, ``R0`` 3 x 3, ``P2`` 3 x 4,
    cutout.py:35-52) and the image shape (rows, cols): M1 = V2C.T @ R0.T exactly as ``lidar_to_rect`` forms it
    (a float32 product, cutout.py:80), P2, height, width, enabled."""
    m1 = np.dot(np.asarray(calib['Tr_velo2cam'], dtype=np.float32).T, np.asarray(calib['R0'], dtype=np.float32).T)
    assert m1.dtype == np.float32 and m1.shape == (4, 3)
    p2 = np.asarray(calib['P2'], dtype=np.float32).reshape(3, 4)
    return np.concatenate([m1.astype(np.float64).ravel(), p2.astype(np.float64).ravel(),
                           [float(img_shape[0]), float(img_shape[1]), 1.0]])


def read_kitti_calib(calib_file):
    """cutout.py:35-52."""
    with open(calib_file) as f:
        lines = f.readlines()
    row = lambda i: np.array(lines[i].strip().split(' ')[1:], dtype=np.float32)
    return {'P2': row(2).reshape(3, 4), 'P3': row(3).reshape(3, 4), 'R0': row(4).reshape(3, 3),
            'Tr_velo2cam': row(5).reshape(3, 4)}


def cut_boxes_batch(frames, boxes, emit, use_drop=None, drop_labels=(), cameras=None, want_index=False):
    """``frames``: list of (xyzi float32 N x 4, labels N); ``boxes[f]``: list of box dictionaries (the reference's
    annotation dictionary) of frame f; ``emit[f][j]``: label the emitted points of box j must carry, ``EMIT_ANY`` or
    ``EMIT_NONE`` (count only); ``use_drop[f][j]``: also drop the labels in ``drop_labels``; ``cameras[f]``: None or a
    ``camera_record``.  Returns ``[[BoxCut, ...] per frame]``."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    nf = len(frames)
    assert nf > 0 and len(boxes) == nf and len(emit) == nf
    offs = np.zeros(nf + 1, dtype=np.int64)
    for i, f in enumerate(frames):
        offs[i + 1] = offs[i] + len(f[0])
    max_points = int(np.max(np.diff(offs)))
    d_xyzi = torch.from_numpy(np.ascontiguousarray(np.concatenate(
        [np.asarray(f[0], dtype=np.float32).reshape(-1, 4) for f in frames]))).cuda()
    d_lab = torch.from_numpy(np.concatenate([np.asarray(f[1]).reshape(-1).astype(np.uint32).view(np.int32) for f in frames])).cuda()
    d_offs = torch.from_numpy(offs).cuda()
    d_cam = None
    if cameras is not None and any(c is not None for c in cameras):
        cam = np.zeros((nf, 27), dtype=np.float64)
        for i, c in enumerate(cameras):
            if c is not None:
                cam[i] = c
        d_cam = torch.from_numpy(cam).cuda()
    drop = np.asarray(list(drop_labels), dtype=np.int32)
    d_drop = torch.from_numpy(drop).cuda() if len(drop) else None
    stream = torch.cuda.current_stream().cuda_stream
    chunks = max(1, (max_points + 4095) // 4096)
    out = [[None] * len(b) for b in boxes]
    most = max(len(b) for b in boxes)
    for r0 in range(0, most, MAX_BOXES_PER_FRAME):          # frames with more than 64 boxes take several rounds
        sel = [list(range(r0, min(len(b), r0 + MAX_BOXES_PER_FRAME))) for b in boxes]
        nb = sum(len(s) for s in sel)
        if nb == 0:
            continue
        box_off = np.zeros(nf + 1, dtype=np.int32)
        box_off[1:] = np.cumsum([len(s) for s in sel])
        recs = np.stack([bx.box_record(boxes[f][j]) for f in range(nf) for j in sel[f]])
        keep = np.array([int(emit[f][j]) for f in range(nf) for j in sel[f]], dtype=np.int32)
        udrop = np.array([int(bool(use_drop[f][j])) if use_drop is not None else 0 for f in range(nf) for j in sel[f]],
                         dtype=np.int32)
        d_box, d_boff = torch.from_numpy(recs).cuda(), torch.from_numpy(box_off).cuda()
        d_keep, d_udrop = torch.from_numpy(keep).cuda(), torch.from_numpy(udrop).cuda()
        d_in = torch.empty(nb, dtype=torch.int32, device="cuda")
        d_fov = torch.empty(nb, dtype=torch.int32, device="cuda")
        d_chunk = torch.empty(nb * chunks, dtype=torch.int32, device="cuda")
        d_ooff = torch.empty(nb + 1, dtype=torch.int64, device="cuda")
        _lib.check(lib.r3d_cut_objects_count(
            d_xyzi.data_ptr(), d_lab.data_ptr(), d_offs.data_ptr(), nf, max_points, d_box.data_ptr(), d_boff.data_ptr(),
            d_keep.data_ptr(), d_udrop.data_ptr(), nb, max(len(s) for s in sel), d_cam.data_ptr() if d_cam is not None else None,
            d_drop.data_ptr() if d_drop is not None else None, len(drop), d_in.data_ptr(), d_fov.data_ptr(),
            d_chunk.data_ptr(), d_ooff.data_ptr(), stream), "cut_objects_count")
        ooff = d_ooff.cpu().numpy()
        total = int(ooff[-1])
        o_xyzi = torch.empty((max(total, 1), 4), dtype=torch.float32, device="cuda")
        o_lab = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
        o_idx = torch.empty(max(total, 1), dtype=torch.int32, device="cuda") if want_index else None
        _lib.check(lib.r3d_cut_objects_write(
            d_xyzi.data_ptr(), d_lab.data_ptr(), d_offs.data_ptr(), nf, max_points, d_box.data_ptr(), d_boff.data_ptr(),
            d_keep.data_ptr(), d_udrop.data_ptr(), nb, d_drop.data_ptr() if d_drop is not None else None, len(drop),
            d_chunk.data_ptr(), d_ooff.data_ptr(), o_xyzi.data_ptr(), o_lab.data_ptr(),
            o_idx.data_ptr() if o_idx is not None else None, stream), "cut_objects_write")
        h_in, h_fov = d_in.cpu().numpy(), d_fov.cpu().numpy()
        h_xyzi, h_lab = o_xyzi.cpu().numpy(), o_lab.cpu().numpy().view(np.uint32)
        h_idx = o_idx.cpu().numpy() if o_idx is not None else None
        k = 0
        for f in range(nf):
            for j in sel[f]:
                a, b = int(ooff[k]), int(ooff[k + 1])
                out[f][j] = BoxCut(int(h_in[k]), int(h_fov[k]), h_xyzi[a:b].copy(), h_lab[a:b].copy(),
                                   h_idx[a:b].copy() if h_idx is not None else None)
                k += 1
    return out
