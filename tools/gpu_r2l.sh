#!/bin/bash
# round 2, GPU call L: k_tail_ws with 12 vs 16 compute warps (rebuild on the box), c2 + c4 per-kernel tables
mkdir -p gpurun_out
for w in 12 16; do
  FDT_NVCC_FLAGS="-DFDT_TAIL_WARPS=$w" python -m face_detection_tflite_b200.build --force > /dev/null 2>&1 || echo build failed
  export FDT_NVCC_FLAGS="-DFDT_TAIL_WARPS=$w"
  FDT_TS=1 FDT_TAIL=1 timeout 300 python tools/tail_check.py 2>&1 | grep -E "oracle|launches" | head -8
  for c in c2 c4; do
    timeout 300 python bench.py --config $c --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_w${w}_$c.json 2> gpurun_out/bench_w${w}_$c.err
    python - <<PY
import json
d=json.load(open('gpurun_out/bench_w${w}_$c.json'))
print("warps $w $c value %.0f"%d["value"], " ".join("%s=%.0f"%(k["kernel"][2:],k["ms"]*1e3) for k in d["kernels"] if k["kernel"] in ("k_tail_ws","k_fc_tc")))
PY
  done
  unset FDT_NVCC_FLAGS
done
