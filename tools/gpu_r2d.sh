#!/bin/bash
# round 2, GPU call D: k_block_ts / k_tail_ws variants vs the layer-by-layer kernels and the f64 oracle; c2 bench per variant
mkdir -p gpurun_out
rm -f gpurun_out/heads_*.npz
for v in "0 0" "1 0" "0 1" "1 1"; do set -- $v
  FDT_TS=$1 FDT_TAIL=$2 timeout 300 python tools/tail_check.py >> gpurun_out/variant_check.log 2>&1
done
cat gpurun_out/variant_check.log
for v in "1 0" "1 1"; do set -- $v
  FDT_TS=$1 FDT_TAIL=$2 timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2_ts$1_tail$2.json 2> gpurun_out/bench_c2_ts$1_tail$2.err; tail -c 300 gpurun_out/bench_c2_ts$1_tail$2.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_ts$1_tail$2.json'))
print("ts$1 tail$2 value %.0f e2e %.0f launches %d faces %d"%(d["value"],d["e2e"]["value"],d["gpu_launches"],d["faces_per_step"]))
for k in d["kernels"]: print("   %-14s %-40s %.3f ms %6.0f GB/s"%(k["kernel"],k["tensor"][:40],k["ms"],k["gbs"] or 0))
PY
done
