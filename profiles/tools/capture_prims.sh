#!/bin/bash
# usage: profiles/tools/capture_prims.sh TAG   (on the GPU box): ncu --set full of the streaming primitives and the per-frame
# preparation kernels, the steady-state launches of the `primitives` section of a full bench run
TAG=$1
CMD="python bench.py --steps 4 --warmup 3 --streams 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_prims_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_prims_$TAG.log; exit 1; }
for k in hadamard_ac_grid_kernel quant_kernel sub_dct_grid_kernel lowres_core_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_$TAG $CMD > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:pixel_cmp_grid_kernel -s 3 -c 1 -f -o gpurun_out/prof_satd8x8_$TAG $CMD > gpurun_out/ncu_satd_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"aq_kernel|lowres_kernel|intra_kernel" -s 40 -c 3 -f -o gpurun_out/prof_prep_$TAG $CMD > gpurun_out/ncu_p_$TAG.log 2>&1
ls -la gpurun_out/prof_*_$TAG.ncu-rep
