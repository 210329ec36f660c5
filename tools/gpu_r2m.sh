#!/bin/bash
# round 2, GPU call M: k_chain_wide (full-range wide layers) vs layer-by-layer kernels and the oracle; c3 bench
mkdir -p gpurun_out
FDT_CHAIN=0 timeout 300 python tools/wide_check.py 2>&1 | tail -6
FDT_CHAIN=1 timeout 300 python tools/wide_check.py 2>&1 | tail -8
timeout 300 python bench.py --config c3 --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_m_c3.json 2> gpurun_out/bench_m_c3.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_m_c3.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_m_c3.json'))
print("c3 value %.0f"%d["value"])
for k in d["kernels"]: print("   %-14s %-44s %.3f ms %6.0f GB/s"%(k["kernel"],k["tensor"][:44],k["ms"],k["gbs"] or 0))
PY
