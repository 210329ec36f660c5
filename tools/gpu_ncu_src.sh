#!/bin/bash
# source-level (CUDA line) stall samples for selected kernels.  $1 = tag
TAG=${1:-x}
mkdir -p gpurun_out
python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_dwpw|k_stem' -s 42 -c 21 -o /tmp/prof_$TAG -f python tools/prof_target.py > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>gpurun_out/ncu3.log
for sel in "k_stem 0" "k_dwpw 0" "k_dwpw 1" "k_dwpw 2" "k_dwpw 12"; do set -- $sel
  ncu -i /tmp/prof_$TAG.ncu-rep --page source --print-source cuda --csv -k regex:$1 --launch-skip $2 --launch-count 1 > gpurun_out/src_${1}_$2_$TAG.csv 2>>gpurun_out/ncu3.log
done
ls -la gpurun_out/ | head -30; tail -n 2 gpurun_out/ncu2.log
