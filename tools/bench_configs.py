"""Device-resident throughput of the two other single-GPU configurations of BASELINE.json (parity cases C2 and C4 of
tests/test_gpu_configs.py; NOT the headline bench): C2 = semantic segmentation, SemanticKITTI-shape scans (124 992
points), 64 x 2048 range image, 20 objects per scan; C4 = OS1-128-shape scans (262 144 points), 128 x 2048 image,
50 objects per scan.  One scan replicated with different schedules, steps dealt to 4 resident engines like bench.py.
Prints one JSON line per configuration.  usage: python tools/bench_configs.py [steps]"""
import json, os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pcl_augmentation_b200 import synth
from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case


def run(name, task, shape, n_obj, rows, cols, n_scans, steps, engines=4):
    base = synth.make_case(task, 7300, shape=shape, number_of_object=n_obj)
    classes = base.config["insertion"]["classes"]
    cases = []
    for j in range(n_scans):
        sched = synth.make_schedule(7400 + j, len(classes), n_obj, [len(base.db[k]) for k in classes])
        c = synth.Case(base.task, base.config, base.pcl5, base.box_lines, base.db, sched, maps=base.maps, cars=base.cars)
        c.pose, c.map_data = base.pose, base.map_data
        cases.append(c)
    inputs = [scan_input_from_case(c) for c in cases]
    engs = [Real3DEngine(task, base.config, base.db, max_scans=n_scans, max_points=len(base.pcl5), rows=rows, cols=cols,
                         max_events=n_obj + 1, max_boxes=128, map_data=base.map_data, sub_batches=2) for _ in range(engines)]
    staged = engs[0].stage(inputs)
    for e in engs:
        e.load(staged); e.run(); e.sync()
    inserted = sum(len(r.inserted) for r in engs[0].unpack(engs[0].fetch_raw())) / n_scans

    def work(w, n):
        for _ in range(w, n, engines):
            engs[w].reset(); engs[w].run()
    def all_steps(n):
        th = [threading.Thread(target=work, args=(w, n)) for w in range(engines)]
        [t.start() for t in th]; [t.join() for t in th]
    all_steps(engines)
    for e in engs:
        e.sync()
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ends = [torch.cuda.Event(enable_timing=True) for _ in engs]
    ev0.record(engs[0].cuda_stream())
    all_steps(steps)
    for ev, e in zip(ends, engs):
        ev.record(e.cuda_stream())
    for e in engs:
        e.sync()
    ms = max(ev0.elapsed_time(ev) for ev in ends)
    for e in engs:
        e.close()
    print(json.dumps({"config": name, "scans_per_step": n_scans, "points_per_scan": len(base.pcl5), "range_image": [rows, cols],
                      "objects_requested": n_obj, "objects_inserted_per_scan": round(inserted, 2), "steps": steps,
                      "resident_engines": engines, "ms_per_step": round(ms / steps, 3),
                      "scans_per_s": round(n_scans * steps / (ms / 1e3), 1)}), flush=True)


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    run("C2 semseg SemanticKITTI-shape 64x2048, 20 objects", "ss", synth.SEMKITTI_SHAPE, 20, 64, 2048, 64, steps)
    run("C4 OS1-128-shape 128x2048, 50 objects", "od", synth.OS128_SHAPE, 50, 128, 2048, 32, steps)
