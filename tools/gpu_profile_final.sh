bash tools/gpu_profile_r2.sh r4d
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_scan_walk -s 1 -c 1 -f -o gpurun_out/r4d_walk_c2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --depth 1 --resident-depth 1 --side-configs= --config c2 > /dev/null 2> gpurun_out/r4d_walk_c2.err
ls -la gpurun_out/r4d_*
