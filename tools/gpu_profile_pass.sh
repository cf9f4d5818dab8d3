#!/bin/bash
# One GPU-box pass: smoke, the side benches, the bench line, the ncu launch list and the full captures of the new kernels.
# usage (through gpurun): bash tools/gpu_profile_pass.sh <tag>
TAG=${1:-r1b}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --side cut_objects > gpurun_out/${TAG}_cut_objects_bench.json 2> gpurun_out/${TAG}_cut_objects_bench.err; tail -1 gpurun_out/${TAG}_cut_objects_bench.json
python bench.py --side rich_map_ss > gpurun_out/${TAG}_rich_map_ss_bench.json 2> gpurun_out/${TAG}_rich_map_ss_bench.err; tail -1 gpurun_out/${TAG}_rich_map_ss_bench.json
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
for d in 12 16; do timeout 300 python bench.py --no-cpu-baseline --resident-depth $d > gpurun_out/${TAG}_bench_d$d.json 2>/dev/null; done
for f in gpurun_out/${TAG}_bench*.json; do python - $f <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["gpu_launches"], round(d["ms_per_step"], 2), round(d["single_batch_ms"], 2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --depth 1 --resident-depth 1 --sub-batches 1 > /dev/null 2>&1
wc -l gpurun_out/${TAG}_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_rms_extents|k_rms_raster|k_rms_finalize" -s 8 -c 3 -f \
    -o gpurun_out/${TAG}_full_richmap_ss python tools/bench_rich_map_ss.py 256 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_cut_count|k_cut_write" -s 6 -c 2 -f \
    -o gpurun_out/${TAG}_full_cutdb python tools/bench_cut_objects.py 64 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
