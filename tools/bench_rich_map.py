"""Rich-map generation (object detection): frames/s of the CUDA operator, host-to-host and device-only.
Prints one JSON line.  usage: python tools/bench_rich_map.py [frames per batch] [steps]
`python bench.py --side rich_map_od` runs the same and adds the CPU baseline (the numpy oracle port on one host core)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pcl_augmentation_b200 import _lib, synth
from pcl_augmentation_b200.object_detection.rich_map import single_drivable_area_map as rm

def gpu_bench(n=256, steps=10):
    """Returns (result dict, CPU-baseline inputs: list of N x 5 float64 clouds)."""
    base = [synth.make_scan(7000 + i, synth.KITTI_SHAPE, synth.make_scene_cars(7000 + i, 6)) for i in range(16)]
    xyzi = [base[i % 16][0] for i in range(n)]
    labels = [base[i % 16][1] & 0xFFFF for i in range(n)]
    for _ in range(3):
        rm.drivable_area_maps_batch(xyzi, labels, 40)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = rm.drivable_area_maps_batch(xyzi, labels, 40)          # host buffers in, host maps out
    dt = time.perf_counter() - t0
    # device-only: the two kernels on resident inputs
    lib = _lib.load()
    offs = np.arange(n + 1, dtype=np.int64) * len(xyzi[0])
    d_x = torch.from_numpy(np.concatenate(xyzi)).cuda(); d_l = torch.from_numpy(np.concatenate(labels).astype(np.int32)).cuda()
    d_o = torch.from_numpy(offs).cuda(); d_d = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib.r3d_rich_map_od_extents(d_x.data_ptr(), d_o.data_ptr(), n, d_d.data_ptr(), st)
    dims = d_d.cpu().numpy(); moff = np.zeros(n + 1, dtype=np.int64); moff[1:] = np.cumsum(dims[:, 0].astype(np.int64) * dims[:, 1])
    total = int(moff[-1]); d_m = torch.from_numpy(moff).cuda()
    road = torch.empty(total, dtype=torch.uint8, device="cuda"); ped = torch.empty_like(road); scr = torch.empty(2 * total, dtype=torch.uint8, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        lib.r3d_rich_map_od_extents(d_x.data_ptr(), d_o.data_ptr(), n, d_d.data_ptr(), st)
        lib.r3d_rich_map_od_build(d_x.data_ptr(), d_l.data_ptr(), d_o.data_ptr(), n, 40, d_d.data_ptr(), d_m.data_ptr(), total,
                                  road.data_ptr(), ped.data_ptr(), scr.data_ptr(), st)
    ev1.record(); torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / steps
    bytes_alg = n * len(xyzi[0]) * 20 * 2                      # both kernels read the 20 B point record once
    res = ({"metric": "rich maps/s (120k-pt KITTI frame: road + pedestrian-area map)", "frames_per_batch": n,
                      "e2e_frames_per_s": round(n * steps / dt, 1), "device_frames_per_s": round(n / (dev_ms / 1e3), 1),
                      "device_ms_per_batch": round(dev_ms, 3), "algorithmic_gbs": round(bytes_alg / (dev_ms / 1e3) / 1e9, 1),
                      "map_cells_per_frame": int(total / n)})
    return res, [np.hstack((xyzi[k], labels[k].reshape(-1, 1))).astype(np.float64) for k in range(min(n, 16))]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print(json.dumps(gpu_bench(n, steps)[0]))

if __name__ == "__main__":
    main()
