#!/bin/bash
# round 2, GPU call W: k_block_ts epilogue with contiguous shares / 16-column units: parity + c2 (default and FDT_TS=2) + c3 + c4
mkdir -p gpurun_out
FDT_TS=1 FDT_TAIL=1 timeout 300 python tools/tail_check.py 2>&1 | grep -E "oracle|launches|Error|error" | head -8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-c2} --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_w_$name.json 2> gpurun_out/bench_w_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_w_$name.json'))
print("$name value %.0f  "%d["value"]+" ".join("%s=%.0f"%(k["kernel"][2:],k["ms"]*1e3) for k in d["kernels"][:14]))
PY
}
run default A=1
run ts2 FDT_TS=2
CFG=c3 run c3 A=1
CFG=c4 run c4 A=1
