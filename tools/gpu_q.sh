#!/bin/bash
# quick GPU iteration: detector parity suite + device-resident bench lines (usage: gpu_q.sh <tag> [configs...])
tag=$1; shift; cfgs=${@:-c2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for c in $cfgs; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${tag}_$c.json'))
print("$c value %.0f  "%d["value"]+" ".join("%s=%.0f"%(k["kernel"][2:],k["ms"]*1e3) for k in d["kernels"][:16]))
PY
done
