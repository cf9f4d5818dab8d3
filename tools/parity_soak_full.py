"""Parity soak at BASELINE sizes: scans of a bench workload (c3: 120 000 points, 10 objects, 1024 yaws, 112 x 1440; c2:
124 992 points, 20 objects, 64 x 2048; c4: 262 144 points, 50 objects, 128 x 2048) through the CUDA engine vs the numpy
oracle.  usage (GPU box): python tools/parity_soak_full.py c3 24"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import real3d_oracle as orc                     # checker only
from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case
from tests.test_gpu_engine import assert_matches_oracle

name, n = sys.argv[1], int(sys.argv[2])
w = bench.WORKLOADS[name]
cases = bench.build_cases(name, 0)[:n]
kw = bench.engine_kwargs(name, cases); kw["max_scans"] = len(cases)
t0 = time.time()
eng = Real3DEngine(w["task"], cases[0].config, cases[0].db, **kw)
res = eng.augment_batch([scan_input_from_case(c) for c in cases])
eng.close()
bad, inserted = [], 0
for i, (c, got) in enumerate(zip(cases, res)):
    ref = orc.augment_scan(c.task, c.pcl5, c.box_lines, c.db, c.schedule.counts, c.schedule.perms, c.config, maps=c.maps,
                           map_data=c.map_data, transform_matrix=c.pose, mode="closed", yaw_steps=w["yaw"], num_row=w["rows"],
                           num_column=w["cols"])
    want = orc.save_arrays(c.task, ref)
    try:
        assert got.status == 0
        assert_matches_oracle(c, got, ref, want)
        inserted += len(got.inserted)
    except AssertionError as exc:
        bad.append((i, int(got.status), str(exc)[:160]))
print(json.dumps({"workload": name, "scans": len(cases), "points_per_scan": len(cases[0].pcl5), "mismatching_scans": bad,
                  "objects_inserted": inserted, "seconds": round(time.time() - t0, 1)}))
