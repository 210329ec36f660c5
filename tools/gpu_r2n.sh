#!/bin/bash
# round 2, GPU call N: whole GPU suite on the wide chains + row-staged letterbox; c2/c3/c4 short benches
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1 ) 2>&1 | grep real; tail -5 gpurun_out/pytest_gpu.log
for c in c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n_${c}.json 2> gpurun_out/bench_n_${c}.err; echo "bench $c rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_n_$c.json'))
print("$c value %.0f e2e %.0f launches %d roofline %s %.3f pcie %s"%(d["value"],d["e2e"]["value"],d["gpu_launches"],d["roofline"]["kernel"],d["roofline"]["frac"],d["e2e"].get("pcie")))
print("   "+" ".join("%s=%.0f"%(k["kernel"][2:],k["ms"]*1e3) for k in d["kernels"]))
PY
done
