#!/bin/bash
mkdir -p gpurun_out
for c in ${CHUNKS:-64 128 256 512}; do
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --chunk $c 2>gpurun_out/bench_q.err | C=$c python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[chunk %s] value %d  sum %.3f ms | '%(os.environ['C'],d['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks))
" || tail -3 gpurun_out/bench_q.err
done
