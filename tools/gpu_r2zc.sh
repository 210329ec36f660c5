#!/bin/bash
# round 2, GPU call Z-c (2 GPUs, final kernels): multi-device handle in one process; bench under torchrun at N=2 (weak + strong) with the PCIe ceiling
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_full_mode.py -m gpu -q -x -k "multi_device" 2>&1 | tail -4
for sc in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --scaling $sc > gpurun_out/bench_z_c2_n2_$sc.json 2> gpurun_out/bench_z_c2_n2_$sc.err; echo "bench $sc rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_z_c2_n2_$sc.json').read().strip().splitlines()[-1])
print("$sc n=%d value %.0f e2e %.0f pcie %s"%(d["n_gpus"],d["value"],d["e2e"]["value"],d["e2e"].get("pcie")))
PY
done
nvidia-smi topo -m 2>&1 | head -12
