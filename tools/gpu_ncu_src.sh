#!/bin/bash
# SASS-level stall samples for selected kernels + launch list.  $1 = tag
TAG=${1:-x}
mkdir -p gpurun_out
python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 42 -c 21 --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_target.py > gpurun_out/ncu1.log 2>&1
python tools/prof_target.py > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_dwpw|k_block|k_stem|k_letterbox|k_decode' -s 42 -c 21 -o /tmp/prof_$TAG -f python tools/prof_target.py > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>gpurun_out/ncu3.log
for sel in "k_stem 0" "k_block 0" "k_block 2" "k_block 12"; do set -- $sel
  ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:$1 --launch-skip $2 --launch-count 1 > gpurun_out/sass_${1}_$2_$TAG.csv 2>>gpurun_out/ncu3.log
done
ls -la gpurun_out/ | head -30; tail -n 2 gpurun_out/ncu2.log
