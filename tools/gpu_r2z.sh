#!/bin/bash
# round 2, GPU call Z: ncu evidence of the final kernels.  c2 (chunk 1024, 10 launches per chunk) and c3 (chunk 256, 39 launches):
# launch list (gpu__time_duration) + full set, each after the same command exited 0 without ncu; c4 mesh net full set
mkdir -p gpurun_out
python tools/prof_target.py 1024 > gpurun_out/prof_plain_c2.log 2>&1 && tail -1 gpurun_out/prof_plain_c2.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 10 --csv --log-file gpurun_out/launches_r02z_c2.csv python tools/prof_target.py 1024 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -s 20 -c 10 -o /tmp/prof_c2 -f python tools/prof_target.py 1024 > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_c2.ncu-rep --page raw --csv > gpurun_out/raw_r02z_c2.csv 2> gpurun_out/ncu3.log
python tools/prof_target.py 256 full > gpurun_out/prof_plain_c3.log 2>&1 && tail -1 gpurun_out/prof_plain_c3.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 78 -c 39 --csv --log-file gpurun_out/launches_r02z_c3.csv python tools/prof_target.py 256 full > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -s 78 -c 39 -o /tmp/prof_c3 -f python tools/prof_target.py 256 full > gpurun_out/ncu5.log 2>&1
ncu -i /tmp/prof_c3.ncu-rep --page raw --csv > gpurun_out/raw_r02z_c3.csv 2>> gpurun_out/ncu3.log
ls -la gpurun_out | grep r02z; tail -n 2 gpurun_out/ncu2.log gpurun_out/ncu5.log
