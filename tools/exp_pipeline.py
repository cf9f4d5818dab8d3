"""Experiment: e2e scans/s of ScanPipeline for several (depth, scans per engine batch) on one GPU."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pcl_augmentation_b200.pipeline import ScanPipeline
from pcl_augmentation_b200.engine import scan_input_from_case

def main():
    torch.cuda.set_device(0)
    total = 256
    cases = bench.build_cases(0, total, bench.DISTINCT_SCANS)
    inputs = [scan_input_from_case(c) for c in cases]
    n_points = len(cases[0].pcl5)
    steps = int(os.environ.get("STEPS", "5"))
    for depth, sub in [(1, 256), (2, 256), (3, 256), (2, 128), (3, 128), (4, 128), (4, 64), (6, 64)]:
        pipe = ScanPipeline("od", cases[0].config, cases[0].db, depth=depth, max_scans=sub, max_points=n_points,
                            rows=bench.ROWS, cols=bench.COLS, yaw_steps=bench.YAW_STEPS, max_events=bench.N_OBJECTS + 1)
        staged = [pipe.engines[0].stage(inputs[i:i + sub]) for i in range(0, total, sub)]
        pipe.warmup(staged[0]); pipe.warmup(staged[-1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = pipe.process(staged * steps)
        dt = time.perf_counter() - t0
        print(json.dumps({"depth": depth, "sub": sub, "steps": steps, "scans_per_s": total * steps / dt,
                          "ms_per_step": 1e3 * dt / steps}), flush=True)
        pipe.close()

if __name__ == "__main__":
    main()
