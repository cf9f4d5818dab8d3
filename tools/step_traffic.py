"""DRAM traffic of ONE whole step from an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch; tools/gpu_profile_r2.sh):

  python tools/step_traffic.py <launches.csv> <out.json> [which step, default: the last complete one]

A step of the serial mode = the launches from one k_reset_state (re-arm from the raw points) to the next k_out_write."""
import csv, json, os, sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
kn, mn, mv, idc, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
launches = {}
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]              # kernel name without namespaces
    launches.setdefault(int(r[idc]), {"kernel": name})[r[mn]] = float(r[mv].replace(",", "")) * SCALE.get(r[mu], 1.0)
seq = [launches[k] for k in sorted(launches)]
starts = [i for i, l in enumerate(seq) if l["kernel"] == "k_reset_state"]
steps = []
for a in starts:
    ends = [i for i in range(a, len(seq)) if seq[i]["kernel"] == "k_out_write"]
    nxt = [s for s in starts if s > a]
    if ends and (not nxt or ends[0] < nxt[0]) and any(seq[i]["kernel"] == "k_scan_walk" for i in range(a, ends[0])):
        steps.append((a, ends[0]))
which = int(sys.argv[3]) if len(sys.argv) > 3 else len(steps) - 1
a, b = steps[which]
per = {}
for l in seq[a:b + 1]:
    e = per.setdefault(l["kernel"], {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    e["launches"] += 1
    e["us"] += l.get("gpu__time_duration.sum", 0.0)
    e["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
tot_us = sum(e["us"] for e in per.values())
tot_b = sum(e["dram_bytes"] for e in per.values())
out = {"dram_bytes_per_step": int(tot_b), "device_us_per_step_serialised": round(tot_us, 1),
       "source": f"profiles/{os.path.basename(src)} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                 f"--clock-control none over `bench.py --steps 1 --warmup 1 --depth 1 --resident-depth 1`; step {which} of {len(steps)}, "
                 f"launches {a}..{b}; per-launch times are cold-cache and serialised)",
       "kernels": {k: {"launches": e["launches"], "us": round(e["us"], 1), "share_of_device_time": round(e["us"] / tot_us, 4),
                       "dram_bytes": int(e["dram_bytes"])} for k, e in sorted(per.items(), key=lambda kv: -kv[1]["us"])}}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
