#!/bin/bash
# compute-sanitizer memcheck + racecheck over the hot path: smoke(), one OD and one semseg engine
# test in both execution models, the rich-map and cut-object kernels.  usage (through gpurun): bash tools/gpu_sanitizer_r2.sh
mkdir -p gpurun_out
run() {   # name tool command...
    local name=$1 tool=$2; shift 2
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 3 "$@" > gpurun_out/r2_sanitizer_${tool}_${name}.full 2>&1
    local rc=$?
    { echo "command: compute-sanitizer --tool $tool $*"; echo "exit code: $rc";
      grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/r2_sanitizer_${tool}_${name}.full | sort | uniq -c | sort -rn | head -40; } \
        > gpurun_out/r2_sanitizer_${tool}_${name}.txt
    tail -3 gpurun_out/r2_sanitizer_${tool}_${name}.txt
}
SMOKE='import __graft_entry__ as g; g.smoke()'
OD='tests/test_gpu_engine.py::test_engine_matches_reference_run[e2e_od_a-walker]'
ODS='tests/test_gpu_engine.py::test_engine_matches_reference_run[e2e_od_a-staged]'
SS='tests/test_gpu_engine.py::test_engine_matches_reference_run[e2e_ss_a-walker]'
SSS='tests/test_gpu_engine.py::test_engine_matches_reference_run[e2e_ss_a-staged]'
for tool in memcheck racecheck; do
    run smoke $tool python -c "$SMOKE"
    run od_walker $tool python -m pytest -x -q "$OD"
    run ss_walker $tool python -m pytest -x -q "$SS"
    run od_staged $tool python -m pytest -x -q "$ODS"
    run ss_staged $tool python -m pytest -x -q "$SSS"
    run richmap $tool python -m pytest -x -q tests/test_gpu_rich_map.py -k "reference_script"
    run cutdb $tool python -m pytest -x -q tests/test_gpu_cut_objects.py -k "script"
done
ls gpurun_out/r2_sanitizer_*.txt
