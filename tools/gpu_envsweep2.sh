#!/bin/bash
# env sweep: ENVS is ';'-separated env assignments
mkdir -p gpurun_out
IFS=';' read -ra SETS <<< "${ENVS:-}"
for e in "${SETS[@]}"; do
  env $e timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS:-} 2>gpurun_out/sweep.err | E="$e" python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[%s] value %d  sum %.3f ms | '%(os.environ['E'],d['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks))
" || tail -3 gpurun_out/sweep.err
done
