"""Cut-object database building: frames/s of the CUDA path (semseg cut_out: every annotated box of a
frame in one pass).  Prints one JSON line.
usage: python tools/bench_cut_objects.py [frames] [steps]
`python bench.py --side cut_objects` runs the same and adds the CPU baseline (the numpy oracle port on one host core)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pcl_augmentation_b200 import _lib, boxes as bx, cut_objects as co, synth
from pcl_augmentation_b200.semantic_segmentation.cut_object import cut_out
from tests.helpers import cut_object_cases


def gpu_bench(n=64, steps=10):
    """Returns (result dict, CPU-baseline inputs: (frames, config))."""
    cases = cut_object_cases("ss", shape=synth.KITTI_SHAPE)
    cfg = cases[0].config
    frames = []
    for i in range(n):
        c = cases[i % len(cases)]
        frames.append((c.pcl5[:, :4].astype(np.float32), c.pcl5[:, 4].astype(np.uint32), [l + "\n" for l in c.box_lines], "00", f"{i:06d}"))
    n_points = sum(len(f[0]) for f in frames)
    n_boxes = sum(1 for f in frames for l in f[2] if int(l.split(" ")[0]) in cfg["insertion"]["classes"])
    for _ in range(2):
        out = cut_out.cut_frames(frames, cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = cut_out.cut_frames(frames, cfg)             # host frames in, host samples out
    dt = time.perf_counter() - t0
    # device-only: both passes on resident inputs
    lib = _lib.load()
    offs = np.zeros(n + 1, dtype=np.int64); offs[1:] = np.cumsum([len(f[0]) for f in frames])
    rows = [[cut_out.line_box(l) for l in f[2] if int(l.split(" ")[0]) in cfg["insertion"]["classes"]] for f in frames]
    boff = np.zeros(n + 1, dtype=np.int32); boff[1:] = np.cumsum([len(r) for r in rows])
    recs = np.stack([bx.box_record(b[0]) for r in rows for b in r])
    keep = np.array([b[0]["class"][0] if isinstance(b[0]["class"], list) else b[0]["class"] for r in rows for b in r], dtype=np.int32)
    nb = len(recs); max_points = int(np.max(np.diff(offs))); chunks = (max_points + 4095) // 4096
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_x, d_l, d_o = d(np.concatenate([f[0] for f in frames])), d(np.concatenate([f[1] for f in frames]).view(np.int32)), d(offs)
    d_b, d_bo, d_k, d_u = d(recs), d(boff), d(keep), d(np.zeros(nb, dtype=np.int32))
    d_in, d_fov = torch.empty(nb, dtype=torch.int32, device="cuda"), torch.empty(nb, dtype=torch.int32, device="cuda")
    d_ch, d_oo = torch.empty(nb * chunks, dtype=torch.int32, device="cuda"), torch.empty(nb + 1, dtype=torch.int64, device="cuda")
    o_x, o_l = torch.empty((n_points, 4), dtype=torch.float32, device="cuda"), torch.empty(n_points, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def both():
        _lib.check(lib.r3d_cut_objects_count(d_x.data_ptr(), d_l.data_ptr(), d_o.data_ptr(), n, max_points, d_b.data_ptr(), d_bo.data_ptr(),
                                             d_k.data_ptr(), d_u.data_ptr(), nb, int(np.max(np.diff(boff))), None, None, 0, d_in.data_ptr(),
                                             d_fov.data_ptr(), d_ch.data_ptr(), d_oo.data_ptr(), st), "count")
        _lib.check(lib.r3d_cut_objects_write(d_x.data_ptr(), d_l.data_ptr(), d_o.data_ptr(), n, max_points, d_b.data_ptr(), d_bo.data_ptr(),
                                             d_k.data_ptr(), d_u.data_ptr(), nb, None, 0, d_ch.data_ptr(), d_oo.data_ptr(), o_x.data_ptr(),
                                             o_l.data_ptr(), None, st), "write")
    both(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        both()
    ev1.record(); torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / steps
    bytes_alg = n_points * (16 + 16)                        # each pass reads xyzi; labels only for the rare hits
    res = ({"metric": "frames/s through cut_out (120k-pt frame, every annotated box in one pass)", "frames": n,
                      "boxes": n_boxes, "samples_saved": sum(len(o) for o in out),
                      "e2e_frames_per_s": round(n * steps / dt, 1), "device_frames_per_s": round(n / (dev_ms / 1e3), 1),
                      "device_ms_per_batch": round(dev_ms, 3), "algorithmic_gbs": round(bytes_alg / (dev_ms / 1e3) / 1e9, 1),
                      })
    return res, (frames[:min(n, 6)], cfg)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print(json.dumps(gpu_bench(n, steps)[0]))


if __name__ == "__main__":
    main()
