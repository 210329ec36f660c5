#!/bin/bash
mkdir -p gpurun_out
for c in ${CHUNKS:-512 1024 2048}; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --chunk $c 2>gpurun_out/bench_q.err | C=$c python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[chunk %s] value %d e2e %d sum %.3f ms | '%(os.environ['C'],d['value'],d['e2e']['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks))
" || tail -3 gpurun_out/bench_q.err
done
