#!/bin/bash
# combined build-flag x env sweep: COMBOS is ';'-separated "NVCCFLAGS|ENV" pairs
mkdir -p gpurun_out
IFS=';' read -ra SETS <<< "${COMBOS:-}"
for c in "${SETS[@]}"; do
  fs="${c%%|*}"; es="${c##*|}"
  FDT_NVCC_FLAGS="$fs" python face_detection_tflite_b200/build.py --force > /dev/null 2>gpurun_out/sweep_build.err || { echo "build failed for [$fs]"; continue; }
  env FDT_NVCC_FLAGS="$fs" $es timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/sweep.err | C="$c" python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[%s] value %d  sum %.3f ms | '%(os.environ['C'],d['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks))
" || tail -3 gpurun_out/sweep.err
done
