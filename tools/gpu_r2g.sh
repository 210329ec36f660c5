#!/bin/bash
# round 2, GPU call G: whole GPU suite, the three bench configs (+ reference arm), ncu launch list of the c2 bench and one --set full capture
mkdir -p gpurun_out
TAG=${1:-r02d}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 180 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for c in c2 c3 c4; do
  T0=$(date +%s); timeout 900 python bench.py --config $c > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; echo "bench $c rc=$? wall=$(( $(date +%s) - T0 ))s"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_$c.json').read().strip().splitlines()[-1])
print("$c value %.0f e2e %.0f launches %d ms/step %.3f roof %s frac %.3f cpu %.0f"%(d["value"],d["e2e"]["value"],d["gpu_launches"],d["ms_per_step"],d["roofline"]["kernel"],d["roofline"]["frac"],(d.get("cpu_baseline") or {}).get("value",0)))
for k in d["kernels"]: print("   %-14s %-40s %.4f ms %6.0f GB/s"%(k["kernel"],k["tensor"][:40],k["ms"],k["gbs"] or 0))
PY
done
T0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 ))s"; cut -c1-300 gpurun_out/bench_${TAG}_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu.log 2>&1; echo "ncu launches rc=$?"
python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 20 -c 10 -o /tmp/prof_$TAG -f python tools/prof_target.py > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>gpurun_out/ncu3.log
for k in k_tail k_block_ws k_block_ts k_stem; do
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:$k --launch-count 1 > gpurun_out/sass_${k}_$TAG.csv 2>>gpurun_out/ncu3.log
done
tail -n 2 gpurun_out/ncu2.log
