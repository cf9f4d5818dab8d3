"""Sequence-wide semseg rich map (semantic_segmentation/rich_map/drivable_area_map.py):
frames/s of the CUDA operator (device-resident and host-to-host).
Prints one JSON line.  usage: python tools/bench_rich_map_ss.py [frames] [steps]
`python bench.py --side rich_map_ss` runs the same and adds the CPU baseline (the numpy oracle port on one host core)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pcl_augmentation_b200 import synth
from pcl_augmentation_b200.semantic_segmentation.rich_map import drivable_area_map as srm


def gpu_bench(n=256, steps=10):
    """Returns (result dict, CPU-baseline inputs: (frames, placement labels))."""
    labels_cfg = synth.load_config("ss")["insertion"]["placement_labels"]
    base = [synth.make_scan(7100 + i, synth.SEMKITTI_SHAPE, synth.make_scene_cars(7100 + i, 6)) for i in range(16)]
    pose0 = synth.make_pose(7100)
    frames = []
    for i in range(n):                                    # a drive: 1.1 m per frame along a slow curve
        a = 0.004 * i
        step = np.eye(4)
        step[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        step[:3, 3] = [1.1 * i, 0.002 * i * i, 0.0]
        frames.append((base[i % 16][0], base[i % 16][1] & 0xFFFF, pose0 @ step))
    n_points = sum(len(f[0]) for f in frames)
    for _ in range(2):
        out = srm.sequence_map(frames, labels_cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = srm.sequence_map(frames, labels_cfg)        # host frames in, host map out
    dt = time.perf_counter() - t0
    # device-only: both passes over resident frames
    b = srm.SequenceMapBuilder(labels_cfg)
    chunk = b.add_extent(frames, keep=False)
    b.begin_raster()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        b.first = True
        b.add_extent(chunk)
        b.keymap.zero_()
        b.order_base = 0
        b.add_raster(chunk)
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / steps
    res = b.finish()
    assert np.array_equal(res["map"], out["map"])
    bytes_alg = n_points * (16 + 20)                      # extent pass reads xyzi, raster pass reads xyzi + label
    res = ({"metric": "frames/s into the sequence map (125k-pt SemanticKITTI frame, two passes)", "frames": n,
                      "e2e_frames_per_s": round(n * steps / dt, 1), "device_frames_per_s": round(n / (dev_ms / 1e3), 1),
                      "device_ms_per_sequence": round(dev_ms, 3), "algorithmic_gbs": round(bytes_alg / (dev_ms / 1e3) / 1e9, 1),
                      "map_shape": list(out["map"].shape),
                      "cells_per_class": [int((out["map"] == v).sum()) for v in (1, 2, 3)]})
    return res, (frames[:min(n, 24)], labels_cfg)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print(json.dumps(gpu_bench(n, steps)[0]))


if __name__ == "__main__":
    main()
