#!/bin/bash
# round 2, GPU call Y: depthwise taps of k_block_ts from the kernel parameters (constant bank), block 2 on k_block_ts: full GPU suite + c2 / c3 / c4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_y.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-c2} --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_y_$name.json 2> gpurun_out/bench_y_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_y_$name.json'))
print("$name value %.0f  "%d["value"]+" ".join("%s=%.0f"%(k["kernel"][2:],k["ms"]*1e3) for k in d["kernels"][:14]))
PY
}
run default A=1
CFG=c3 run c3 A=1
CFG=c4 run c4 A=1
