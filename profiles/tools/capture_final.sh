#!/bin/bash
# usage: [X264CU_LIB=...] capture_final.sh TAG : ncu --set full of one steady-state decision-kernel launch
TAG=$1
CMD="python bench.py --steps 4 --warmup 3 --streams 1 --no-cpu-baseline --no-primitives --value-only"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:final_kernel_ -s 20 -c 1 -f -o gpurun_out/prof_final_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
ls -la gpurun_out/prof_final_$TAG.ncu-rep
