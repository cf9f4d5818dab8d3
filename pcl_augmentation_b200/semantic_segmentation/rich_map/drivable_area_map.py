"""Sequence-wide rich map of the semantic-segmentation pipeline on the GPU — the reference's
``semantic_segmentation/rich_map/drivable_area_map.py`` (``ss/rm``): every frame of a sequence is moved to the world
frame, a 1 m grid spans the xy extent of all frames and the road / parking / sidewalk points are rasterised in order
(1 = road, 2 = parking, 3 = sidewalk, sticky; ss/rm:122-206).  Output ``.npz{map: float64 X x Y, move: int 4 x 1}``.

    python -m pcl_augmentation_b200.semantic_segmentation.rich_map.drivable_area_map --sequence 00 \
        [--config ../config/semantic-kitti.yaml]

``sequence_map`` is the operator (frames in, map out); ``SequenceMapBuilder`` feeds the frames in chunks.
"""
import os

import numpy as np

from ... import _lib


def surface_table(placement_labels):
    """(labels, classes) with the precedence of ss/rm:192-200: class 1 is tested first, then 3, the rest of the
    surface labels (ss/rm:104-105) is class 2."""
    table = {}
    for lab in placement_labels[2]:
        table[int(lab)] = 2
    for lab in placement_labels[3]:
        table[int(lab)] = 3
    for lab in placement_labels[1]:
        table[int(lab)] = 1
    labs = sorted(table)
    return np.array(labs, dtype=np.int32), np.array([table[l] for l in labs], dtype=np.int32)


class _Chunk:
    """Frames of one call, packed and resident on the device."""

    def __init__(self, frames):
        import torch
        n = len(frames)
        offs = np.zeros(n + 1, dtype=np.int64)
        for i, f in enumerate(frames):
            offs[i + 1] = offs[i] + len(f[0])
        self.n, self.total = n, int(offs[-1])
        self.max_points = int(np.max(np.diff(offs))) if n else 0
        xyzi = np.concatenate([np.asarray(f[0], dtype=np.float32).reshape(-1, 4) for f in frames])
        labels = np.concatenate([np.asarray(f[1]).reshape(-1).astype(np.uint32).view(np.int32) for f in frames])
        poses = np.stack([np.asarray(f[2], dtype=np.float64).reshape(4, 4) for f in frames])
        self.xyzi = torch.from_numpy(np.ascontiguousarray(xyzi)).cuda()
        self.labels = torch.from_numpy(labels).cuda()
        self.offs = torch.from_numpy(offs).cuda()
        self.poses = torch.from_numpy(np.ascontiguousarray(poses)).cuda()


class SequenceMapBuilder:
    """Two passes like the reference (extent of all frames, then the raster), frames fed in chunks:
    ``add_extent(frames)`` for every chunk, ``begin_raster()``, ``add_raster(frames)`` for every chunk in the same
    order, ``finish()``.  ``frames`` = list of (xyzi float32 N x 4, labels N, lidar->world 4 x 4).  Chunks passed to
    ``add_extent(..., keep=True)`` stay on the device and ``raster_kept()`` replays them."""

    def __init__(self, placement_labels):
        import torch
        _lib.require_cuda()
        self.lib = _lib.load()
        self.torch = torch
        labs, cls = surface_table(placement_labels)
        self.n_surface = len(labs)
        self.d_labs = torch.from_numpy(labs).cuda()
        self.d_cls = torch.from_numpy(cls).cuda()
        self.ext = torch.zeros(4, dtype=torch.int64, device="cuda")
        self.out5 = torch.zeros(5, dtype=torch.int64, device="cuda")
        self.err = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.first = True
        self.kept = []
        self.keymap = None
        self.order_base = 0

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def add_extent(self, frames, keep=False):
        c = frames if isinstance(frames, _Chunk) else _Chunk(frames)
        _lib.check(self.lib.r3d_rich_map_ss_extents(c.xyzi.data_ptr(), c.offs.data_ptr(), c.poses.data_ptr(), c.n,
                                                    c.max_points, 1 if self.first else 0, self.ext.data_ptr(),
                                                    self.out5.data_ptr(), self._stream()), "rich_map_ss_extents")
        self.first = False
        if keep:
            self.kept.append(c)
        return c

    def begin_raster(self):
        min_x, min_y, size_x, size_y, any_point = (int(v) for v in self.out5.cpu().numpy())
        if not any_point:
            raise ValueError("rich map of a sequence without points")
        self.min_x, self.min_y, self.size_x, self.size_y = min_x, min_y, size_x, size_y
        self.keymap = self.torch.zeros(size_x * size_y, dtype=self.torch.int64, device="cuda")
        self.order_base = 0

    def add_raster(self, frames):
        c = frames if isinstance(frames, _Chunk) else _Chunk(frames)
        _lib.check(self.lib.r3d_rich_map_ss_raster(c.xyzi.data_ptr(), c.labels.data_ptr(), c.offs.data_ptr(),
                                                   c.poses.data_ptr(), c.n, c.max_points, self.d_labs.data_ptr(),
                                                   self.d_cls.data_ptr(), self.n_surface, self.min_x, self.min_y,
                                                   self.size_x, self.size_y, self.order_base, self.keymap.data_ptr(),
                                                   self.err.data_ptr(), self._stream()), "rich_map_ss_raster")
        self.order_base += c.total

    def raster_kept(self):
        for c in self.kept:
            self.add_raster(c)
        self.kept = []

    def finish(self):
        cells = self.size_x * self.size_y
        out = self.torch.empty(cells, dtype=self.torch.uint8, device="cuda")
        _lib.check(self.lib.r3d_rich_map_ss_finalize(self.keymap.data_ptr(), cells, out.data_ptr(), self._stream()),
                   "rich_map_ss_finalize")
        err = int(self.err.item())
        if err == 1:                                              # ss/rm:190
            raise AssertionError("Indexing error: a surface point lies below the map origin")
        if err == 2:
            raise IndexError("a surface point lies outside the sequence map")
        grid = out.cpu().numpy().reshape(self.size_x, self.size_y).astype(np.float64)     # np.zeros((size_x, size_y)), ss/rm:170
        return {'map': grid, 'move': np.array([[self.min_x], [self.min_y], [0], [1]])}


def sequence_map(frames, placement_labels, chunk_frames=256):
    """``frames``: list of (xyzi float32 N x 4, labels N, lidar->world 4 x 4) of ONE sequence in dataset order.
    Returns ``{'map': float64 X x Y in {0, 1, 2, 3}, 'move': int 4 x 1 [[min_x], [min_y], [0], [1]]}``."""
    b = SequenceMapBuilder(placement_labels)
    for i in range(0, len(frames), chunk_frames):
        b.add_extent(frames[i:i + chunk_frames], keep=True)
    b.begin_raster()
    b.raster_kept()
    return b.finish()


def generate_map(config, sequence, chunk_frames=256, device_budget_bytes=64 << 30, log=print):
    """The reference script (ss/rm:91-209) for one SemanticKITTI sequence: writes
    ``<save_path>/maps/small/npz/<sequence>.npz``; the frames stay on the device between the two passes while they
    fit in ``device_budget_bytes`` (20 B per point), otherwise they are read twice like the reference does."""
    from ..Real3DAug.tools.datasets import SemanticKITTI
    dataset = SemanticKITTI(config, sequence)
    save_path = config['path']['maps_path'].split('/')            # ss/rm:97-102: '<root>/maps/small/npz' -> '<root>'
    save_path.pop()
    save_path.pop()
    save_path.pop()
    save_path = '/'.join(save_path)
    os.makedirs(f'{save_path}/maps/small/npz', exist_ok=True)
    b = SequenceMapBuilder(config['insertion']['placement_labels'])
    n = len(dataset)
    read = lambda i0: [dataset.read_frame(i)[:3] for i in range(i0, min(i0 + chunk_frames, n))]
    used, spilled = 0, []
    for i0 in range(0, n, chunk_frames):
        frames = read(i0)
        nbytes = sum(len(f[0]) for f in frames) * 20
        keep = used + nbytes <= device_budget_bytes and not spilled
        b.add_extent(frames, keep=keep)
        if keep:
            used += nbytes
        else:
            spilled.append(i0)
        log(f'extent {min(i0 + chunk_frames, n)} / {n} frames')
    b.begin_raster()
    b.raster_kept()
    for i0 in spilled:
        b.add_raster(read(i0))
    out = b.finish()
    np.savez(f'{save_path}/maps/small/npz/{sequence}', move=out['move'], map=out['map'])
    return out


def main(argv=None):
    import argparse
    import yaml
    ap = argparse.ArgumentParser(description="sequence-wide road / parking / sidewalk map (SemanticKITTI) on the GPU")
    ap.add_argument("--config", default="../config/semantic-kitti.yaml")
    ap.add_argument("--sequence", required=True)
    ap.add_argument("--chunk", type=int, default=256)
    args = ap.parse_args(argv)
    with open(args.config, "r") as f:
        generate_map(yaml.safe_load(f), str(args.sequence).zfill(2), args.chunk)


if __name__ == "__main__":
    main()
