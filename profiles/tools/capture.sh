#!/bin/bash
# usage: profiles/tools/capture.sh TAG   (run on the GPU box through gpurun)
# 1) the plain command must exit 0; 2) launch list; 3) one --set full capture of a steady-state launch of each hot kernel
TAG=$1
CMD="python bench.py --steps 4 --warmup 3 --streams 1 --no-cpu-baseline --no-primitives --value-only"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel_t -s 40 -c 2 -f -o gpurun_out/prof_search_$TAG $CMD > gpurun_out/ncu_s_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:final_kernel_ -s 20 -c 1 -f -o gpurun_out/prof_final_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG*
