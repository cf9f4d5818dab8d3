#!/bin/bash
# A/B of the row-clipped collision walk (R3D_CLIP_MIN_COLS): whole GPU suite with the clip forced on every collision test,
# c2 / c3 bench lines per variant, full-size parity soak of the semseg workload.  Build the variants first (CPU box):
#   bash tools/build_variant.sh base -DR3D_CLIP_MIN_COLS=0; bash tools/build_variant.sh clip1 -DR3D_CLIP_MIN_COLS=1
#   bash tools/build_variant.sh clip4 -DR3D_CLIP_MIN_COLS=4; bash tools/build_variant.sh clip8 -DR3D_CLIP_MIN_COLS=8
# usage: gpurun -- bash tools/gpu_clip_ab.sh   (record: profiles/r2_clip_ab.json)
mkdir -p gpurun_out
V=build_variants/libreal3d_b200
R3D_LIB_PATH=${V}_clip1.so timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r4a_tests_clip1.log 2>&1; echo "tests clip1 rc=$?"; tail -2 gpurun_out/r4a_tests_clip1.log
timeout 200 python tools/parity_soak_full.py c2 8 > gpurun_out/r4a_soak_c2.json 2> gpurun_out/r4a_soak_c2.err; echo "soak rc=$?"; cat gpurun_out/r4a_soak_c2.json
for v in base clip8 clip4; do
  lib=${V}_$v.so
  R3D_LIB_PATH=$lib timeout 200 python bench.py --config c2 --side-configs '' --no-cpu-baseline > gpurun_out/r4a_c2_$v.json 2> gpurun_out/r4a_c2_$v.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r4a_c2_$v.json').read().strip().splitlines()[-1])
print('c2 $v', round(d['value']), d['ms_per_step'], d.get('single_batch_ms'), d['e2e']['value'], d['engine_stats'].get('walker_detail'))
P
done
for v in base clip8 clip1; do
  lib=${V}_$v.so
  R3D_LIB_PATH=$lib timeout 200 python bench.py --side-configs '' --no-cpu-baseline > gpurun_out/r4a_c3_$v.json 2> gpurun_out/r4a_c3_$v.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r4a_c3_$v.json').read().strip().splitlines()[-1])
print('c3 $v', round(d['value']), d['ms_per_step'], d.get('single_batch_ms'), d['e2e']['value'])
P
done
