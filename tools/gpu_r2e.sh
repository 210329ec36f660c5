#!/bin/bash
# round 2, GPU call E: default plan (k_block_ts on) and the reworked tail kernel vs the references; c2 bench per variant
mkdir -p gpurun_out
rm -f gpurun_out/heads_*.npz gpurun_out/variant_check.log
for v in "0 0" "1 0" "2 0" "1 1"; do set -- $v
  FDT_TS=$1 FDT_TAIL=$2 timeout 300 python tools/tail_check.py >> gpurun_out/variant_check.log 2>&1
done
cat gpurun_out/variant_check.log
for v in "1 0 512" "1 1 512" "1 1 592"; do set -- $v
  FDT_TS=$1 FDT_TAIL=$2 timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu --chunk $3 > gpurun_out/bench_c2_ts$1_tail$2_$3.json 2> gpurun_out/bench_c2_ts$1_tail$2_$3.err; tail -c 300 gpurun_out/bench_c2_ts$1_tail$2_$3.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_ts$1_tail$2_$3.json'))
print("ts$1 tail$2 chunk$3 value %.0f e2e %.0f launches %d faces %d"%(d["value"],d["e2e"]["value"],d["gpu_launches"],d["faces_per_step"]))
for k in d["kernels"]: print("   %-14s %-40s %.3f ms %6.0f GB/s"%(k["kernel"],k["tensor"][:40],k["ms"],k["gbs"] or 0))
PY
done
