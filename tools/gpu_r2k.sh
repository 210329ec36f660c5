#!/bin/bash
# round 2, GPU call K: verification of HEAD after the container was re-created: whole GPU suite, smoke, c2/c3/c4 + reference bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for c in c2 c3 c4; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_${c}.json 2> gpurun_out/bench_${c}.err; echo "bench $c rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$c.json'))
print("$c value %.0f e2e %.0f launches %d roofline %s %.3f cpu %s"%(d["value"],d["e2e"]["value"],d["gpu_launches"],d["roofline"]["kernel"],d["roofline"]["frac"],d.get("cpu_baseline",{}).get("value")))
for k in d["kernels"]: print("   %-14s %-44s %.3f ms %6.0f GB/s"%(k["kernel"],k["tensor"][:44],k["ms"],k["gbs"] or 0))
PY
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-400
