"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's rich-map generation.

Pinned by ``tests/golden/rich_map_od.npz`` and ``tests/golden/rich_map_ss.npz`` (written by ``oracle/make_golden.py``
from the UNMODIFIED ``object_detection/rich_map/single_drivable_area_map.py``, abbreviated ``od/rm`` below, and
``semantic_segmentation/rich_map/drivable_area_map.py``, ``ss/rm``).  Only ``tests/`` and the CPU
baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np


def disk(radius):
    """skimage.morphology.disk (od/rm:8,151,183): cells with dx^2 + dy^2 <= r^2 of the (2r+1)^2 square."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return (X ** 2 + Y ** 2) <= radius ** 2


def _dilate(img, fp):
    """Binary dilation, cells outside the image ignored (what skimage / scipy 'reflect' give for a symmetric convex
    footprint: a mirrored cell always lies inside the clipped window)."""
    r = fp.shape[0] // 2
    out = np.zeros_like(img)
    H, W = img.shape
    for dr in range(-r, r + 1):
        for dc in range(-r, r + 1):
            if not fp[dr + r, dc + r]:
                continue
            src = img[max(0, dr):H + min(0, dr), max(0, dc):W + min(0, dc)]
            out[max(0, -dr):H + min(0, -dr), max(0, -dc):W + min(0, -dc)] |= src
    return out


def _erode(img, fp):
    r = fp.shape[0] // 2
    out = np.ones_like(img)
    H, W = img.shape
    for dr in range(-r, r + 1):
        for dc in range(-r, r + 1):
            if not fp[dr + r, dc + r]:
                continue
            src = img[max(0, dr):H + min(0, dr), max(0, dc):W + min(0, dc)]
            out[max(0, -dr):H + min(0, -dr), max(0, -dc):W + min(0, -dc)] &= src
    return out


def rich_map_od(point_cloud, road_label):
    """od/rm:123-194 for one frame.  ``point_cloud`` N x 5 (x, y, z, intensity, label) as ``KITTI.__getitem__`` returns
    it.  Returns (road_map uint8 X x Y, pedestrian_map uint8 X x Y, min_x, min_y)."""
    x, y = point_cloud[:, 0], point_cloud[:, 1]
    min_x, min_y = int(min(x)), int(min(y))                       # od/rm:123-124: int() truncates toward zero
    max_x, max_y = int(max(x)) + 1, int(max(y)) + 1               # od/rm:126-127
    size_x, size_y = int(max_x - min_x), int(max_y - min_y)
    road = point_cloud[:, 4] == road_label                        # od/rm:137
    ix = (x[road] - min_x).astype(np.int64)                       # od/rm:140-145: int() of a float64 difference
    iy = (y[road] - min_y).astype(np.int64)                       # (negative indices wrap like Python's)
    raster = np.zeros((size_x, size_y), dtype=bool)
    raster[ix, iy] = True
    closed = _erode(_dilate(raster, disk(4)), disk(4))            # od/rm:150-156 closing(disk(4))
    near = _dilate(closed, np.ones((3, 3), dtype=bool))           # od/rm:164-180: 8-neighbourhood of the road
    ring = near & ~closed
    ped = _dilate(ring, disk(2))                                  # od/rm:182-188 dilation(disk(2))
    return closed.astype(np.uint8), ped.astype(np.uint8), min_x, min_y


def surface_classes(placement_labels):
    """label -> map class with the precedence of ss/rm:192-200 (``ss/rm`` =
    ``semantic_segmentation/rich_map/drivable_area_map.py``): class 1 is tested first, then 3, everything else in the
    surface list is 2."""
    table = {}
    for lab in placement_labels[2]:
        table[int(lab)] = 2
    for lab in placement_labels[3]:
        table[int(lab)] = 3
    for lab in placement_labels[1]:
        table[int(lab)] = 1
    return table


def rich_map_ss(frames, placement_labels):
    """ss/rm:122-206 for one sequence.  ``frames``: list of (pcl N x 5 float64 = x, y, z, intensity, label;
    4 x 4 lidar->world) in dataset order.  Returns ``{'map': float64 X x Y in {0, 1, 2, 3}, 'move': int 4 x 1}``.

    Per cell the reference's in-order writes (ss/rm:192-200: 1 / 2 unless the cell holds 3; 3 unconditionally) end as
    3 if any class-3 point fell into the cell, else as the class of the LAST class-1/2 point that did."""
    table = surface_classes(placement_labels)
    world = []
    for pcl, t_matrix in frames:
        pts = np.array(pcl[:, :4], dtype=np.float64)
        pts[:, 3] = 1                                            # ss/rm:135
        pts = t_matrix @ pts.T                                   # ss/rm:136 (BLAS dgemm: a fused multiply-add chain)
        world.append((pts / pts[3, :]).T)                        # ss/rm:137
    allp = np.concatenate(world)
    min_x, min_y = int(np.floor(allp[:, 0].min())), int(np.floor(allp[:, 1].min()))      # ss/rm:158-159
    max_x, max_y = int(allp[:, 0].max()) + 1, int(allp[:, 1].max()) + 1                  # ss/rm:161-162
    size_x, size_y = int(max_x - min_x), int(max_y - min_y)
    labels = np.concatenate([np.asarray(pcl[:, 4]) for pcl, _ in frames]).astype(np.int64)
    cls = np.zeros(len(labels), dtype=np.int64)
    for lab, c in table.items():
        cls[labels == lab] = c
    keep = cls > 0
    px, py = allp[keep, 0] - min_x, allp[keep, 1] - min_y        # ss/rm:188-189
    assert not ((px < 0) | (py < 0)).any()                       # ss/rm:190
    flat = px.astype(np.int64) * size_y + py.astype(np.int64)    # int(): truncation
    c = cls[keep]
    order = np.arange(1, len(c) + 1, dtype=np.int64)
    sticky = np.zeros(size_x * size_y, dtype=bool)
    sticky[flat[c == 3]] = True
    key = np.zeros(size_x * size_y, dtype=np.int64)
    m = c != 3
    np.maximum.at(key, flat[m], order[m] * 4 + c[m])
    grid = np.where(sticky, 3, key & 3).astype(np.float64).reshape(size_x, size_y)
    return {'map': grid, 'move': np.array([[min_x], [min_y], [0], [1]])}
