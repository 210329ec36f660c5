#!/bin/bash
# per-instruction executed counts / stall samples of one k_block_ts launch (variant $1, skip $2 launches of k_block_ts)
mkdir -p gpurun_out
v=$1; skip=${2:-12}
export FDT_CUDA_LIB=$PWD/variants/lib_$v.so
python tools/prof_target.py 1024 2>&1 | tail -1
ncu --clock-control none --section SourceCounters --section WarpStateStats --import-source on -k regex:k_block_ts -s $skip -c 1 -o /tmp/src_$v -f python tools/prof_target.py 1024 > gpurun_out/ncu_src_$v.log 2>&1
ncu -i /tmp/src_$v.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$v.csv 2>> gpurun_out/ncu_src_$v.log
