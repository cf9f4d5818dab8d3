"""Diagnosis: phase timeline of ScanPipeline (depth 3, 256 scans per batch)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pcl_augmentation_b200.pipeline import ScanPipeline
from pcl_augmentation_b200.engine import scan_input_from_case

torch.cuda.set_device(0)
total = 256
cases = bench.build_cases(0, total, bench.DISTINCT_SCANS)
inputs = [scan_input_from_case(c) for c in cases]
n_points = len(cases[0].pcl5)
for depth, excl in ((1, False), (2, True), (3, True), (2, False), (3, False)):
    pipe = ScanPipeline("od", cases[0].config, cases[0].db, depth=depth, exclusive_run=excl, max_scans=total, max_points=n_points,
                        rows=bench.ROWS, cols=bench.COLS, yaw_steps=bench.YAW_STEPS, max_events=bench.N_OBJECTS + 1)
    staged = pipe.engines[0].stage(inputs)
    pipe.warmup(staged); pipe.warmup(staged)
    torch.cuda.synchronize()
    trace = []
    t0 = time.perf_counter()
    pipe.process([staged] * 8, trace=trace)
    dt = time.perf_counter() - t0
    print("depth", depth, excl, "ms/step", 1e3 * dt / 8)
    for w, i, a, b, c, d in sorted(trace, key=lambda r: r[2]):
        print(f"  w{w} batch{i}: start {1e3*(a-t0):7.1f}  load {1e3*(b-a):6.1f}  run {1e3*(c-b):6.1f}  fetch {1e3*(d-c):6.1f}")
    pipe.close()
