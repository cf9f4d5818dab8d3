"""Summarise `ncu --set full` captures into small tracked files under profiles/:

  python tools/ncu_summary.py <round tag> <rep> [<rep> ...]

writes profiles/<tag>_ncu_<rep name>.csv (one row per captured launch, the metrics the roofline discussion uses) and
merges dram traffic per launch into profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum"]
# ncu kernel name -> bench.py kernel-table name
BENCH_NAME = {"k_project": "project_zbuffer_full", "k_close_fill": "close_fill_full", "k_clear_images": "clear_images_full",
              "k_minmax": "minmax_elevation_full", "k_ingest": "ingest_spherical", "k_update": "update_mask_patch",
              "k_out_write": "compact_output", "k_out_count": "compact_output", "k_onmap": "placement",
              "k_close_fill_raw_pipelined": "close_fill_full", "k_close_fill_tasks": "close_fill_full", "k_close_fill_tma": "close_fill_full",
              "k_onmap_full": "placement", "k_road_level": "placement", "k_collide": "placement",
              "k_occl_count": "occlusion_count", "k_select_emit": "select_emit",
              # round 2: the per-scan walker and the fused streaming passes
              "k_scan_walk": "scan_walk", "k_ingest_count": "ingest_spherical", "k_scatter_project": "scatter_project",
              "k_bucket_scan3": "index_build", "k_grid_near_tiled": "index_build", "k_walk_prepare": "walk_prepare"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    args = [a for a in sys.argv[1:] if a != "--no-traffic"]
    keep_traffic = "--no-traffic" in sys.argv          # captures of another workload than the headline one: CSV only
    tag, reps = args[0], args[1:]
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, units, data = rows[0], rows[1], rows[2:]
        metrics = [m for m in METRICS if m in h]
        idx = [h.index(m) for m in metrics]
        kn = h.index("Kernel Name")
        name = os.path.splitext(os.path.basename(rep))[0]
        dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_{name.replace(tag + '_', '')}.csv")
        agg, extra = {}, {}
        with open(dst, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel"] + [f"{m} [{units[i]}]" for m, i in zip(metrics, idx)])
            for r in data:
                k = r[kn].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]        # kernel name without namespaces
                w.writerow([k] + [r[i] for i in idx])
                rd = float(r[idx[1]].replace(",", "")) * SCALE.get(units[idx[1]], 1.0)
                wr = float(r[idx[2]].replace(",", "")) * SCALE.get(units[idx[2]], 1.0)
                b = BENCH_NAME.get(k)
                if b:
                    agg.setdefault(b, []).append((k, rd + wr))
                    def pct(metric):
                        return round(float(r[h.index(metric)].replace(",", "")), 2) if metric in h else None
                    extra.setdefault(b, {}).setdefault(k, {"sm_throughput_pct": pct("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                                                           "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                                           "warps_active_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active"),
                                                           "ncu_duration_us": round(float(r[idx[0]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[idx[0]], 1.0), 1)})
        for b, items in agg.items():
            per_kernel = {}
            for k, v in items:
                per_kernel.setdefault(k, []).append(v)
            total = sum(sum(v) / len(v) for v in per_kernel.values())      # mean per launch, summed over the stage's kernels
            longest = max(extra[b].values(), key=lambda v: v["ncu_duration_us"])       # the stage's dominant kernel
            traffic[b] = {"bytes_per_launch": round(total), "kernels": sorted(per_kernel), **longest,
                          "source": f"profiles/{os.path.basename(dst)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                                    f"all scans of the batch active in the captured launch)"}
        print("wrote", dst)
    if keep_traffic:
        return
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print("wrote", traffic_path)


if __name__ == "__main__":
    main()
