"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's rich-map generation.

Pinned by ``tests/golden/rich_map_od.npz`` (written by ``oracle/make_golden.py`` from the UNMODIFIED
``object_detection/rich_map/single_drivable_area_map.py``; abbreviated ``od/rm`` below).  Only ``tests/`` and the CPU
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
