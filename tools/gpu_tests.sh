#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
for c in 256 512 1024 2048; do timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --chunk $c 2>gpurun_out/bench_c$c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk',$c,'value',round(d['value']),'ms/step',round(d['ms_per_step'],2), 'sum kernels ms', round(sum(k['ms'] for k in d['kernels']),3))
"; done
