#!/bin/bash
# Final pass of the round on the GPU box: sanitizer over the walker tests of both pipelines, the GPU suite, the default
# bench line.  usage: gpurun -- bash tools/gpu_final_r2.sh <tag>
TAG=${1:-r4b}
mkdir -p gpurun_out
for tool in racecheck memcheck; do
  for t in e2e_ss_a e2e_od_a; do
    timeout 300 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 3 python -m pytest -x -q \
      "tests/test_gpu_engine.py::test_engine_matches_reference_run[$t-walker]" > gpurun_out/${TAG}_san_${tool}_$t.full 2>&1
    echo "$tool $t rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_san_${tool}_$t.full | tail -2
  done
done
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
