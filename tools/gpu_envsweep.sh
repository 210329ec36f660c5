#!/bin/bash
# runtime-env sweep: ENVSETS is a ';'-separated list of "VAR=val VAR2=val" sets
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -1
IFS=';' read -ra SETS <<< "${ENVSETS:-}"
for es in "${SETS[@]}"; do
  env $es timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS:-} 2>gpurun_out/sweep.err | ES="$es" python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[%s] value %d  sum_kernels %.3f ms | '%(os.environ['ES'],d['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks))
" || tail -3 gpurun_out/sweep.err
done
