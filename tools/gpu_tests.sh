#!/bin/bash
mkdir -p gpurun_out
timeout 90 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; rc=$?; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; fi
timeout 400 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -${TAILN:-25} gpurun_out/pytest_gpu.log
for c in ${CHUNKS:-256}; do timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --chunk $c 2>gpurun_out/bench_c$c.err > gpurun_out/bench_c$c.json; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c$c.json').read().strip().splitlines()[-1])
    print('chunk',$c,'value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']) if d.get('e2e') else None,'h2d',d['e2e']['h2d_bytes_per_step'] if d.get('e2e') else None)
    tot=sum(k['ms'] for k in d['kernels'])
    for k in d['kernels']: print("  %2d %-12s %-18s %.4f ms %5.1f%% %6.2f TF %7.1f GB/s"%(k['launch'],k['kernel'],k['tensor'][:18],k['ms'],100*k['ms']/tot,k['tflops'] or 0,k['gbs'] or 0))
    print('  sum',round(tot,3))
except Exception as e:
    print('bench failed',e); print(open('gpurun_out/bench_c$c.err').read()[-2000:])
PY
done
