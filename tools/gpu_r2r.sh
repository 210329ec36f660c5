#!/bin/bash
# round 2, GPU call R: quick parity (tail_check: short-range / back vs f64 oracle) + c2 / c4 benches
mkdir -p gpurun_out
FDT_TS=1 FDT_TAIL=1 timeout 300 python tools/tail_check.py 2>&1 | grep -E "oracle|launches|Error|error" | head -8
for c in ${CFGS:-c2 c4}; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r_$c.json 2> gpurun_out/bench_r_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_r_$c.json'))
print("$c value %.0f  "%d["value"]+" ".join("%s=%.0f"%(k["kernel"][2:],k["ms"]*1e3) for k in d["kernels"]))
PY
done
