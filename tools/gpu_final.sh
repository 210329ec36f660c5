#!/bin/bash
# End-of-round style run: smoke, GPU tests, reference arm, default bench (with CPU baseline).
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
timeout 180 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo skip-pytest
T0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 ))s"; cut -c1-500 gpurun_out/bench_reference.json
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
k=d.pop('kernels'); print(json.dumps(d,indent=1)[:3000])
PY
