"""bench.py — augmented scans/sec of the Real3D-Aug hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the one the metric is quoted on): a batch of 256 synthetic KITTI-shape HDL-64
scans (120 000 points each, 112 x 1440 range image), 10 cut pedestrians / cyclists to insert per scan, 1024 yaw
candidates per cut object, sharded by scan: every rank owns its own batch of 256 (weak scaling, no collective on
the data path).  One step = one pass of the whole hot path (placement search + occlusion + insertion + output
compaction) over the rank's batch.

  value  : scans/s with the batch already resident in HBM (device re-arm + run, CUDA events on the engine stream)
  e2e    : scans/s through the public API (ScanPipeline) with HOST (pinned) buffers: H2D of every scan + run + D2H of
           the results, consecutive batches overlapped on 3 engines / streams
  roofline: the kernel with the largest share of device time; achieved = algorithmic bytes / measured time
  cpu_baseline: the numpy oracle port of the reference timed on one host core on a bounded sample (rank 0, N = 1)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCANS_PER_GPU = 256
DISTINCT_SCANS = 32
N_OBJECTS = 10
YAW_STEPS = 1024
ROWS, COLS = 112, 1440
METRIC = "augmented scans/sec (120k-pt KITTI scan, placement + occlusion + insertion)"


def workload_config(n_gpus):
    return {"workload": "batch of 256 KITTI-shape scans per GPU (120000 pts, 112x1440 range image), 10 cut "
                        "pedestrians/cyclists per scan, 1024 yaw candidates per object, sharded by scan",
            "scans_per_gpu": SCANS_PER_GPU, "points_per_scan": 120000, "objects_per_scan": N_OBJECTS,
            "yaw_candidates": YAW_STEPS, "range_image": [ROWS, COLS], "parallelism": f"scan-sharded x{n_gpus}",
            "l2": "inputs larger than L2 (614 MB of points per batch, no flush needed)",
            "distinct_scans": DISTINCT_SCANS}


def build_cases(rank, n_scans=SCANS_PER_GPU, distinct=DISTINCT_SCANS):
    """`distinct` different synthetic scans per rank, tiled to `n_scans` with different schedules / object draws."""
    from pcl_augmentation_b200 import synth
    base = [synth.make_case("od", 9000 + rank * 1000 + i, number_of_object=N_OBJECTS) for i in range(distinct)]
    cases = []
    for j in range(n_scans):
        c = base[j % distinct]
        sched = synth.make_schedule(50000 + rank * 10000 + j, len(c.config["insertion"]["classes"]), N_OBJECTS,
                                    [len(c.db[k]) for k in c.config["insertion"]["classes"]])
        cases.append(synth.Case(c.task, c.config, c.pcl5, c.box_lines, c.db, sched, maps=c.maps, cars=c.cars))
    return cases


class ClockSampler:
    """SM clock / throttle reasons of this rank's GPU, sampled through NVML every 10 ms from a background thread
    while the timed regions run (the quantities of the profiling recipe's nvidia-smi line: clocks.sm, clocks.max.sm,
    power.draw, clocks_event_reasons.*).  `mark(True/False)` brackets the timed regions; only samples taken inside
    them are summarised."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"),
               ("hw_power_brake", "nvmlClocksEventReasonHwPowerBrakeSlowdown"))

    def __init__(self, cuda_index, period_s=0.01):
        self.period, self.samples, self.inside, self.stop_flag, self.thread = period_s, [], False, False, None
        self.nv, self.handle, self.error = None, None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            self.nv = pynvml
        except Exception as exc:                      # reported in the JSON line, never silently dropped
            self.error = f"NVML unavailable: {exc!r}"

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def mark(self, inside):
        self.inside = inside

    def _loop(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((self.inside, sm, mask, pw))
            except Exception as exc:
                self.error = f"NVML sample failed: {exc!r}"
                return
            time.sleep(self.period)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.error or "NVML unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        nv = self.nv
        mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        timed = [s for s in self.samples if s[0]]
        reasons = set()
        for _, _, mask, _ in timed:
            for name, attr in self.REASONS:
                if mask & getattr(nv, attr, 0):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(s[1] for s in timed) if timed else None, "sm_max_mhz": float(mx),
               "sm_min_mhz": min(s[1] for s in timed) if timed else None,
               "power_w_max": round(max(s[3] for s in timed), 1) if timed else None,
               "reasons": sorted(reasons), "samples": len(timed), "samples_total": len(self.samples),
               "source": "NVML, 10 ms period, samples inside the two timed regions (device-resident + e2e)"}
        if self.error:
            out["error"] = self.error
        return out


# algorithmic bytes per unit: N points, HW pixels, 20 B point record.  "units" = how many scans /
# tries the kernel really processed in the timed region (engine counters), so gated-off launches add time but no bytes.
PLACEMENT_STAGES = ("onmap", "road_level", "collide", "place_try")


def algorithmic_bytes(kernel, n_points, hw, stats, n_scans, steps):
    per_project = stats["projected_scans"]
    per_try = stats["tried_objects"]
    per_mask = stats["masked_scans"]
    table = {
        "project_zbuffer_full": (20 * n_points + 8 * hw, per_project), # z-buffer pass, round 0 of every run
        "clear_images_full": (8 * hw, per_project),
        "close_fill_full": (16 * hw, per_project),
        "update_mask_patch": (5 * n_points, per_mask),                 # occlusion mask
        "placement": (20 * n_points, per_try),                         # placement pass over the scene, all stages
        "compact_output": (40 * n_points, n_scans * steps),
        "ingest_spherical": (20 * n_points, 0),
    }
    per_unit, units = table.get(kernel, (0, 0))
    return per_unit, units


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(kernel)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline_sample(yaw_steps=YAW_STEPS, budget_s=25.0):
    """The numpy oracle port of the reference (oracle/real3d_oracle.py) on ONE host core on whole scans of the same
    workload until ~budget_s of CPU work is spent (at least one scan)."""
    from oracle import real3d_oracle as orc            # checker / baseline only
    from pcl_augmentation_b200 import synth
    t_total, n = 0.0, 0
    while t_total < budget_s and n < 4:
        case = synth.make_case("od", 9000 + n, number_of_object=N_OBJECTS)
        t0 = time.perf_counter()
        orc.augment_scan("od", case.pcl5, case.box_lines, case.db, case.schedule.counts, case.schedule.perms,
                         case.config, maps=case.maps, mode="closed", yaw_steps=yaw_steps, num_row=ROWS, num_column=COLS)
        t_total += time.perf_counter() - t0
        n += 1
    return n / t_total, n, t_total


def _ref_worker(seed):
    from oracle import real3d_oracle as orc
    from pcl_augmentation_b200 import synth
    case = synth.make_case("od", seed, number_of_object=N_OBJECTS)
    t0 = time.perf_counter()
    orc.augment_scan("od", case.pcl5, case.box_lines, case.db, case.schedule.counts, case.schedule.perms, case.config,
                     maps=case.maps, mode="closed", yaw_steps=YAW_STEPS, num_row=ROWS, num_column=COLS)
    return time.perf_counter() - t0


def run_reference(args, json_out):
    """--impl reference: the CPU implementation of the path (the oracle port: the Python reference itself cannot
    travel to the GPU box) on all host cores, one scan per worker process per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 32))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_ref_worker, [9000 + i for i in range(cores)])
            break                                            # one warm-up wave is enough for a CPU path
        t0 = time.perf_counter()
        done = 0
        for s in range(args.steps):
            pool.map(_ref_worker, [9100 + s * cores + i for i in range(cores)])
            done += cores
        dt = time.perf_counter() - t0
    value = done / dt
    sample = f"{cores} whole scans per step (one per worker process), {args.steps} steps"
    print(file=json_out, flush=True, *[json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})])


def run_side(args, json_out):
    """GPU numbers of an offline tool (tools/bench_*.py) + its CPU baseline (bench.py is the only
    place besides tests/ that executes oracle/)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    if args.side == "rich_map_od":
        import bench_rich_map as tool
        from oracle import rich_map_oracle as rmo                  # CPU baseline leg only
        res, clouds = tool.gpu_bench(256, 10)
        t0, k = time.perf_counter(), 0
        while time.perf_counter() - t0 < 5.0:
            rmo.rich_map_od(clouds[k % len(clouds)], 40)
            k += 1
        unit, cpu, sample = "frames/s", k / (time.perf_counter() - t0), f"{k} frames"
    elif args.side == "rich_map_ss":
        import bench_rich_map_ss as tool
        from oracle import rich_map_oracle as rmo                  # CPU baseline leg only
        res, (frames, labels_cfg) = tool.gpu_bench(256, 10)
        t0 = time.perf_counter()
        rmo.rich_map_ss([(np.hstack((p.astype(np.float64), l.reshape(-1, 1).astype(np.float64))), T) for p, l, T in frames],
                        labels_cfg)
        unit, cpu, sample = "frames/s", len(frames) / (time.perf_counter() - t0), f"a {len(frames)}-frame sequence"
    else:
        import bench_cut_objects as tool
        from oracle import cut_objects_oracle as coo               # CPU baseline leg only
        res, (frames, cfg) = tool.gpu_bench(64, 10)
        t0 = time.perf_counter()
        for f in frames:
            coo.cut_objects_ss(np.hstack((f[0].astype(np.float64), f[1].reshape(-1, 1).astype(np.float64))), f[2], cfg, f[3], f[4])
        unit, cpu, sample = "frames/s", len(frames) / (time.perf_counter() - t0), f"{len(frames)} frames"
    res["cpu_baseline"] = {"value": round(cpu, 2), "unit": unit, "cores": 1, "kind": "port", "sample": sample}
    print(json.dumps(res), file=json_out, flush=True)


def main():
    # Libraries (NCCL's version banner, for one) write to fd 1: keep the real stdout for the ONE JSON line and send
    # everything else written to fd 1 during the run to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scans", type=int, default=SCANS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sub-batches", type=int, default=1,
                    help="sub-batches the engine advances concurrently (own streams) in the device-resident leg")
    ap.add_argument("--resident-depth", type=int, default=8,
                    help="engines (each with its own HBM-resident batch) the device-resident leg deals the steps to: "
                         "the thinning last rounds of one step overlap the busy first rounds of the next")
    ap.add_argument("--e2e-sub-batches", type=int, default=2, help="same, per engine of the e2e pipeline")
    ap.add_argument("--depth", type=int, default=4, help="engines (streams) the e2e leg pipelines batches through")
    ap.add_argument("--side", default=None, choices=["rich_map_od", "rich_map_ss", "cut_objects"],
                    help="instead of the headline bench: one of the offline tools either side of the path (SURVEY 8f rows "
                         "3-4), GPU numbers from tools/bench_*.py plus the CPU baseline (numpy oracle port, one host core)")
    args = ap.parse_args()
    if args.side:
        run_side(args, json_out)
        return
    if args.impl == "reference":
        run_reference(args, json_out)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    from pcl_augmentation_b200 import sharding as _sh
    numa_cpus = _sh.bind_to_gpu_numa_node(local_rank) if (world > 1 and os.environ.get("R3D_NUMA_BIND", "1") != "0") else None
    if world > 1:
        # NCCL_DEBUG output goes to fd 1, which main() redirected to stderr: the JSON line keeps the real stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from pcl_augmentation_b200 import sharding

    def max_over_ranks(x):
        return sharding.all_reduce_scalar(x, "max", "cuda")

    def sum_over_ranks(x):
        return sharding.all_reduce_scalar(x, "sum", "cuda")

    from pcl_augmentation_b200.engine import scan_input_from_case
    from pcl_augmentation_b200.pipeline import ScanPipeline
    n_scans = args.scans
    cases = build_cases(rank, n_scans, min(DISTINCT_SCANS, n_scans))
    n_points = len(cases[0].pcl5)
    # PIPE_DEPTH engines (own stream, own device-resident batch, own host thread): the e2e leg streams batches through
    # all of them so H2D, compute and D2H of consecutive batches overlap; the device-resident leg uses the first one
    pipe = ScanPipeline("od", cases[0].config, cases[0].db, depth=args.depth, max_scans=n_scans, max_points=n_points,
                        rows=ROWS, cols=COLS, yaw_steps=YAW_STEPS, max_events=N_OBJECTS + 1, sub_batches=args.sub_batches)
    eng = pipe.engines[0]
    staged = eng.stage([scan_input_from_case(c) for c in cases])
    stream = eng.cuda_stream()
    res_depth = max(1, args.resident_depth)
    from pcl_augmentation_b200.engine import Real3DEngine
    extra_engines = [Real3DEngine("od", cases[0].config, cases[0].db, max_scans=n_scans, max_points=n_points, rows=ROWS,
                                  cols=COLS, yaw_steps=YAW_STEPS, max_events=N_OBJECTS + 1, sub_batches=args.sub_batches)
                     for _ in range(res_depth - args.depth)]
    res_engines = (pipe.engines + extra_engines)[:res_depth]

    # ---- device-resident throughput ("value") --------------------------------------------------------
    # Every resident engine holds the batch in HBM before the timed region starts.  A step = re-arm + all rounds +
    # output compaction of one 256-scan batch; the K steps are dealt round-robin to `res_depth` engines (own host
    # thread, own streams), so consecutive steps overlap: a step alone is bound by the chain of dependent kernels of
    # its longest-running scan (rounds x kernel latency), not by the device.
    for e in res_engines:
        e.set_sub_batches(args.sub_batches)
        e.load(staged)
        e.sync()

    def resident_steps(n_steps):
        errors = []

        def worker(w):
            try:
                for _ in range(w, n_steps, res_depth):
                    res_engines[w].reset(from_raw_points=True)      # the whole device path: ingest + indices included
                    res_engines[w].run()
            except BaseException as exc:
                errors.append(exc)
        threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in range(res_depth)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    resident_steps(max(args.warmup, res_depth))
    for e in res_engines:
        e.sync()
    # one step alone (no overlap between steps): the latency of a batch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(3):
        eng.reset(from_raw_points=True); eng.run()
    ev1.record(stream)
    eng.sync()
    single_batch_ms = ev0.elapsed_time(ev1) / 3
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = eng.launch_count()
    ext_streams = [e.cuda_stream() for e in res_engines]
    ev0 = torch.cuda.Event(enable_timing=True)
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in res_engines]
    barrier()
    sampler.mark(True)
    ev0.record(ext_streams[0])              # every engine stream is idle here (barrier = device synchronize)
    resident_steps(args.steps)
    for ev, es in zip(ev_end, ext_streams):
        ev.record(es)
    for e in res_engines:
        e.sync()
    sampler.mark(False)
    barrier()
    dev_ms = max(ev0.elapsed_time(ev) for ev in ev_end)
    launches = eng.launch_count() - launches0
    # per-kernel CUDA-event times: a second, untimed pass over the same steps with the engine's event profiling on
    # (two event records per launch would otherwise sit inside the timed region)
    # and ONE sub-batch, so that kernels run strictly one after the other and an event pair times one kernel alone
    eng.set_sub_batches(1)
    eng.profile(True)
    for _ in range(args.steps):
        eng.reset(from_raw_points=True)
        eng.run()
    eng.sync()
    prof = eng.profile_read()
    stats = eng.stats()
    eng.profile(False)
    eng.set_sub_batches(args.sub_batches)
    dev_ms = max_over_ranks(dev_ms)
    value = world * n_scans * args.steps / (dev_ms / 1000.0)

    # ---- end to end through the public API with host buffers ("e2e") -----------------------------------
    # every step = one staged batch in pinned host memory: H2D of all points / labels / maps / schedules, the
    # spherical ingest, the augmentation rounds, D2H of the augmented clouds into pinned host buffers
    for e in pipe.engines:
        e.set_sub_batches(args.e2e_sub_batches)
    pipe.warmup(staged)
    pipe.warmup(staged)
    barrier()
    sampler.mark(True)
    t0 = time.perf_counter()
    out_bytes = pipe.process([staged] * args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.mark(False)
    clocks = sampler.stop()
    d2h = sum(out_bytes)
    e2e_value = world * n_scans * args.steps / e2e_s
    h2d = staged["xyzi"].nbytes + staged["labels"].nbytes + staged["boxes"].nbytes + staged["maps"].nbytes + staged["perms"].nbytes
    results = eng.unpack(pipe._buffers[0])
    inserted_total = sum(len(r.inserted) for r in results)

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peak, peak_src = peaks()
    hw = ROWS * COLS
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    # the three placement stages (A5-A10) share ONE figure in SURVEY §8d (20 N bytes per tried object): one group
    group = {"ms": sum(v["ms"] for k, v in prof.items() if k in PLACEMENT_STAGES),
             "launches": max([v["launches"] for k, v in prof.items() if k in PLACEMENT_STAGES] or [0])}
    entries = dict(prof)
    entries["placement"] = group
    kernel_table = {}
    for name, v in sorted(entries.items(), key=lambda kv: -kv[1]["ms"]):
        per_unit, units = algorithmic_bytes(name, n_points, hw, stats, n_scans, args.steps)
        entry = {"ms": round(v["ms"], 3), "launches": v["launches"], "share": round(v["ms"] / max(total_kernel_ms, 1e-9), 4)}
        if name == "placement":
            entry["stages"] = [k for k in PLACEMENT_STAGES if k in prof and prof[k]["launches"]]
        if per_unit and units and v["ms"] > 0:
            gbs = per_unit * units / (v["ms"] / 1000.0) / 1e9
            entry.update(algorithmic_bytes_per_unit=per_unit, units=units, achieved_gbs=round(gbs, 1),
                         frac=round(gbs / peak, 4))
        kernel_table[name] = entry
    roofline = None
    for name, e in kernel_table.items():        # dominant kernel (stage) = largest share of device time with a bytes model
        if "achieved_gbs" in e:
            traffic = measured_traffic(name)
            roofline = {"kernel": name, "bound": "hbm", "achieved": e["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": e["frac"], "traffic": traffic["bytes_per_launch"] if traffic else None,
                        "traffic_source": traffic["source"] if traffic else None, "peak_source": peak_src,
                        "bytes_per_launch": e["algorithmic_bytes_per_unit"] * e["units"] / max(e["launches"], 1),
                        "avg_launch_ms": e["ms"] / max(e["launches"], 1), "share_of_kernel_time": e["share"]}
            break
    b_scan = N_OBJECTS * (45 * n_points + 24 * hw) + 40 * n_points
    step_roofline = {"algorithmic_bytes_per_scan": b_scan, "achieved_gbs": round(value / world * b_scan / 1e9, 1),
                     "frac_of_peak": round(value / world * b_scan / 1e9 / peak, 4)}

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, n_done, secs = cpu_baseline_sample()
        cpu = {"value": v, "unit": "scans/s", "cores": 1, "kind": "port",
               "sample": f"{n_done} whole scan(s) of the same workload through oracle/real3d_oracle.py, {secs:.1f} s",
               "host_cpus": os.cpu_count()}
    inserted_all = sum_over_ranks(inserted_total)
    if rank == 0:
        print(file=json_out, flush=True, *[json.dumps({
            "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(world), steps_in_flight=f"{res_depth} (one 256-scan step per resident engine; "
                           f"`single_batch_ms` is a step alone)"),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h / args.steps), "pipeline_depth": args.depth,
                    "sub_batches_per_engine": args.e2e_sub_batches,
                    "ms_per_step": 1000.0 * e2e_s / args.steps,
                    "pcie_gbs_each_way": round(max(h2d, d2h / args.steps) / (e2e_s / args.steps) / 1e9, 1)},
            "gpu_launches": int(launches), "clocks": clocks, "sub_batches": args.sub_batches,
            "resident_engines": res_depth, "single_batch_ms": single_batch_ms, "numa_bound_cpus": numa_cpus,
            "single_batch_scans_per_s": world * n_scans / (single_batch_ms / 1000.0),
            "kernel_times": "CUDA events around every launch in a separate pass with one sub-batch (serial)",
            "step_roofline": step_roofline, "kernels": kernel_table,
            "rounds_per_step": results[0].extra["rounds"], "objects_inserted_per_scan": inserted_all / (world * n_scans),
            "engine_stats": stats})])
    pipe.close()
    for e in extra_engines:
        e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
