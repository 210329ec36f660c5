#!/bin/bash
# Full-set ncu capture (with SASS stall samples) of the first N k_block_ws launches of one chunk.  $1 = tag, $2 = count
TAG=${1:-x}; CNT=${2:-6}
mkdir -p gpurun_out
python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_block|k_stem' -s ${SKIP:-22} -c $CNT -o /tmp/prof_$TAG -f python tools/prof_target.py > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>gpurun_out/ncu3.log
for i in $(seq 0 $((CNT-1))); do
  ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/sass_${i}_$TAG.csv 2>>gpurun_out/ncu3.log
done
tail -n 2 gpurun_out/ncu2.log
