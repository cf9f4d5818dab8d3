"""Experiment: e2e scans/s of ScanPipeline for several (depth, tail fraction) on one GPU."""
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
    steps = int(os.environ.get("STEPS", "20"))
    for depth, tail, sub in [(3, 8, 1), (3, 8, 2), (3, 8, 3), (3, 256, 2), (3, 16, 2), (4, 8, 2), (3, 4, 2)]:
        pipe = ScanPipeline("od", cases[0].config, cases[0].db, depth=depth, tail_fraction=tail, max_scans=total, max_points=n_points,
                            rows=bench.ROWS, cols=bench.COLS, yaw_steps=bench.YAW_STEPS, max_events=bench.N_OBJECTS + 1, sub_batches=sub)
        staged = pipe.engines[0].stage(inputs)
        pipe.warmup(staged); pipe.warmup(staged)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.process([staged] * steps)
        dt = time.perf_counter() - t0
        print(json.dumps({"depth": depth, "tail_fraction": tail, "sub": sub, "steps": steps, "scans_per_s": round(total * steps / dt),
                          "ms_per_step": round(1e3 * dt / steps, 2)}), flush=True)
        pipe.close()

if __name__ == "__main__":
    main()
