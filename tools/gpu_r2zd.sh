#!/bin/bash
# round 2, GPU call Z-d (8 GPUs, final kernels): bench under torchrun at N=8 (weak) with the all-ranks PCIe ceiling; topology
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_z_c2_n$N.json 2> gpurun_out/bench_z_c2_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_z_c2_n$N.json').read().strip().splitlines()[-1])
print("n=%d value %.0f e2e %.0f pcie %s"%(d["n_gpus"],d["value"],d["e2e"]["value"],d["e2e"].get("pcie")))
PY
nvidia-smi topo -m 2>&1 | head -12 | cut -c1-160
nproc; free -g | head -2
