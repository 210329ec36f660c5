#!/bin/bash
# smoke + short bench with per-kernel table (no pytest)
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu ${BENCH_ARGS:-} 2>gpurun_out/bench_q.err > gpurun_out/bench_q.json; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
    ks=d['kernels']
    print('value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']) if d.get('e2e') else None, 'sum %.3f'%sum(k['ms'] for k in ks))
    print(' '.join('%.0f'%(k['ms']*1e3) for k in ks))
except Exception as e:
    print('bench failed',e); print(open('gpurun_out/bench_q.err').read()[-2000:])
PY
