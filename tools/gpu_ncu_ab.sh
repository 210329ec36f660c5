#!/bin/bash
# ncu A/B of k_block_ts variants (variants/lib_*.so): launches 22, 23 = blocks 1, 2 of the third 1024-frame chunk
mkdir -p gpurun_out
for v in "$@"; do
  export FDT_CUDA_LIB=$PWD/variants/lib_$v.so
  python tools/prof_target.py 1024 2>&1 | tail -1
  ncu --clock-control none --section WarpStateStats --section SchedulerStats --section InstructionStats --section SpeedOfLight --section MemoryWorkloadAnalysis \
      --metrics sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,gpu__time_duration.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
      -k regex:k_block_ts -s 12 -c 2 -o /tmp/ab_$v -f python tools/prof_target.py 1024 > gpurun_out/ncu_ab_$v.log 2>&1
  ncu -i /tmp/ab_$v.ncu-rep --page raw --csv > gpurun_out/ab_$v.csv 2>> gpurun_out/ncu_ab_$v.log
done
