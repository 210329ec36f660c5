#!/bin/bash
# round 2, GPU call H: the generalised k_tail_ws (detector tails + face-landmark chain): parity subset, smoke, c2 / c4 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or full_mode or materialised or variants or detections_match or raw_heads" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 180 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for c in c2 c4; do
  timeout 900 python bench.py --config $c --no-cpu > gpurun_out/bench_r02e_$c.json 2> gpurun_out/bench_r02e_$c.err; echo "bench $c rc=$?"; tail -c 400 gpurun_out/bench_r02e_$c.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02e_$c.json').read().strip().splitlines()[-1])
print("$c value %.0f e2e %.0f launches %d ms/step %.3f roof %s frac %.3f"%(d["value"],d["e2e"]["value"],d["gpu_launches"],d["ms_per_step"],d["roofline"]["kernel"],d["roofline"]["frac"]))
for k in d["kernels"]: print("   %-14s %-40s %.4f ms %6.0f GB/s"%(k["kernel"],k["tensor"][:40],k["ms"],k["gbs"] or 0))
PY
done
