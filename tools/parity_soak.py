"""Parity soak: many fresh seeded scans through the CUDA engine vs the numpy oracle (keep-masks, placement choices,
counts exact; inserted xyz within 1e-6 m) — beyond the fixed seeds of tests/.  usage (GPU box):
    python tools/parity_soak.py od 48 [first_seed]     /    python tools/parity_soak.py ss 12
Uses the oracle as the checker only (like tests/)."""
import json, sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from pcl_augmentation_b200 import synth
from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case
from tests.helpers import GOLDEN_SHAPE
from tests.test_gpu_engine import assert_matches_oracle, oracle_run

task, n = sys.argv[1], int(sys.argv[2])
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
rng = np.random.default_rng(seed0)
t0, bad, inserted, removed = time.time(), [], 0, 0
cases = []
for i in range(n):
    ncls = 2 if task == "od" else 6
    counts = rng.integers(0, 3, ncls).tolist()
    if sum(counts) == 0:
        counts[0] = 1
    cases.append(synth.make_case(task, seed0 + i, shape=GOLDEN_SHAPE, counts=counts, n_cars=int(rng.integers(2, 14)),
                                 obj_range=(4.0, 16.0)))          # one cut-object DB for all scans of an engine
for j0 in range(0, n, 8):                                  # batches of 8 scans with different schedules
    group = cases[j0:j0 + 8]
    eng = Real3DEngine(task, group[0].config, group[0].db, max_scans=len(group), max_points=max(len(c.pcl5) for c in group),
                       map_data=group[0].map_data, max_events=14)
    if task == "ss":                                       # one engine = one sequence map
        for c in group[1:]:
            c.pose, c.map_data = group[0].pose, group[0].map_data
    res = eng.augment_batch([scan_input_from_case(c) for c in group])
    eng.close()
    for ci, (c, got) in enumerate(zip(group, res)):
        ref, want = oracle_run(c)
        try:
            assert got.status == 0
            assert_matches_oracle(c, got, ref, want)
            inserted += len(got.inserted)
            removed += len(c.pcl5) - int(ref["keep_orig"].sum())
        except AssertionError as exc:
            bad.append((seed0 + j0 + ci, int(got.status), [(n_, int(r_)) for n_, r_, _ in got.inserted][:3],
                        [(n_, int(r_)) for n_, r_, _ in ref['inserted']][:3], str(exc)[:120]))
print(json.dumps({"task": task, "scans": n, "first_seed": seed0, "mismatching_scans": bad, "objects_inserted": inserted,
                  "scene_points_removed": removed, "seconds": round(time.time() - t0, 1)}))
