#!/bin/bash
mkdir -p gpurun_out
for c in 128 512; do for d in 511 0; do
  FDT_STEM_DBG=$d timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --chunk $c 2>gpurun_out/sweep.err | D="$d c$c" python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[dbg %s] value %d  sum %.3f ms | '%(os.environ['D'],d['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks[:6]))
" || tail -3 gpurun_out/sweep.err
done; done
