"""Diagnostic: spread of the walker CTA lifetimes over the scans of one batch of a bench workload.
usage (GPU box): python tools/walk_spread.py c2"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cases = bench.build_cases(name, 0)
w = bench.WORKLOADS[name]
eng = Real3DEngine(w["task"], cases[0].config, cases[0].db, **bench.engine_kwargs(name, cases))
staged = eng.stage([scan_input_from_case(c) for c in cases])
eng.load(staged); eng.run(); eng.sync()
eng.reset(from_raw_points=True); eng.run(); eng.sync()
cyc, tries = eng.walk_profile()
res = eng.unpack(eng.fetch_raw())
ms = cyc / 1.965e6
order = np.argsort(ms)
print(name, "scans", len(ms), "mean ms", ms.mean().round(3), "p50", np.median(ms).round(3), "p90", np.quantile(ms, 0.9).round(3),
      "max", ms.max().round(3), "| tries mean", tries.mean().round(2), "max", tries.max())
print("ms per try: mean", (ms / np.maximum(tries, 1)).mean().round(4), "of the slowest 8:", (ms[order[-8:]] / tries[order[-8:]]).round(4))
print("slowest 8 scans:", [(int(i), int(i) % w["distinct"], round(float(ms[i]), 2), int(tries[i]), len(res[i].inserted)) for i in order[-8:]])
print("by base scan (index % distinct): mean ms", [round(float(ms[j::w["distinct"]].mean()), 2) for j in range(min(w["distinct"], 8))])
hist, edges = np.histogram(tries, bins=[0, 11, 13, 16, 21, 26, 31, 41, 61, 101, 1000])
print("tries histogram", list(zip(edges[:-1].tolist(), hist.tolist())))
