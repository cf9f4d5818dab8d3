"""Hot CUDA source lines of a kernel in an .ncu-rep (compile with -lineinfo, capture with --import-source on).
usage: python tools/ncu_source_hot.py <rep> <kernel regex> [top N] [launch index]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split into launches: a launch restarts the file list; detect by repeated (file, function) pairs
tables, cur_file, cur_fn, hdr = [], None, None, None
seen = set()
launch = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        key = (cur_file, cur_fn)
        if key in seen:
            launch += 1; seen = set()
        seen.add(key); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0] != "" and len(r) == len(hdr):
        tables.append((launch, cur_file, cur_fn, hdr, r))
sel = [t for t in tables if t[0] == which]
if not sel:
    print("no rows"); sys.exit(1)
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
tot = 0.0
items = []
for _, f, fn, h, r in sel:
    i_samp = h.index("# Samples"); i_inst = h.index("Instructions Executed")
    stall_idx = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    s = num(r[i_samp]); tot += s
    stalls = sorted(((num(r[i]), h[i][6:]) for i in stall_idx), reverse=True)[:3]
    items.append((s, f.split("/")[-1], r[0], r[1].strip(), r[i_inst], [(n, int(v)) for v, n in stalls if v > 0]))
print(f"launch {which}: total samples {tot:.0f}")
for s, f, ln, src, inst, st in sorted(items, key=lambda x: -x[0])[:top]:
    print(f"{100*s/max(tot,1):5.1f}% {f}:{ln:>4} inst={inst:>9} {src[:100]}  {st}")
