#!/bin/bash
# build-flag sweep on the GPU box: FLAGSETS is a ';'-separated list of nvcc flag sets
mkdir -p gpurun_out
IFS=';' read -ra SETS <<< "${FLAGSETS:-}"
for fs in "${SETS[@]}"; do
  FDT_NVCC_FLAGS="$fs" python face_detection_tflite_b200/build.py --force > /dev/null 2>gpurun_out/sweep_build.err || { echo "build failed for [$fs]"; tail -5 gpurun_out/sweep_build.err; continue; }
  FDT_NVCC_FLAGS="$fs" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS:-} 2>gpurun_out/sweep.err | FS="$fs" python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[%s] value %d  sum_kernels %.3f ms | '%(os.environ['FS'],d['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks))
" || tail -3 gpurun_out/sweep.err
done
