#!/bin/bash
# One GPU-box pass of round 2: the ncu launch list of one whole step (serial mode) and the `--set full` captures of the
# kernels of the step.  usage (through gpurun): bash tools/gpu_profile_r2.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --depth 1 --resident-depth 1 --side-configs="
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > /dev/null 2>&1
wc -l gpurun_out/${TAG}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_scan_walk|k_ingest_count|k_scatter_project|k_bucket_scan3|k_grid_near_bits|k_close_fill|k_out_count|k_out_write" \
    -s 11 -c 11 -f -o gpurun_out/${TAG}_full_step $BENCH > /dev/null 2> gpurun_out/${TAG}_ncu.err
tail -2 gpurun_out/${TAG}_ncu.err
ls -la gpurun_out/${TAG}_*.ncu-rep
