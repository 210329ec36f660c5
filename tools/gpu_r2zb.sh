#!/bin/bash
# round 2, GPU call Z-b: end-of-round style run on the final kernels: smoke, reference arm, default bench (c2, with CPU baseline and e2e), c3, c4
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
timeout 180 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
T0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_z_reference.json 2> gpurun_out/bench_z_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 ))s"; cut -c1-300 gpurun_out/bench_z_reference.json
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_z_c2.json 2> gpurun_out/bench_z_c2.err; echo "bench c2 rc=$? wall=$(( $(date +%s) - T0 ))s"; tail -2 gpurun_out/bench_z_c2.err
for c in c3 c4; do
  T0=$(date +%s); timeout 900 python bench.py --config $c --no-cpu > gpurun_out/bench_z_$c.json 2> gpurun_out/bench_z_$c.err; echo "bench $c rc=$? wall=$(( $(date +%s) - T0 ))s"
done
python - <<'PY'
import json
for c in ("c2","c3","c4"):
    d=json.loads(open('gpurun_out/bench_z_%s.json'%c).read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print(c,"value %.0f"%d["value"],"e2e",e.get("value"),"roofline",r["kernel"],"%.3f"%r["frac"],"traffic",r["traffic"],"cpu",(d.get("cpu_baseline") or {}).get("value"),"launches",d["gpu_launches"],"clocks",d["clocks"])
PY
