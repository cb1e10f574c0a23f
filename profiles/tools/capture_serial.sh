CMD="python bench.py --steps 4 --warmup 3 --streams 1 --no-cpu-baseline --no-primitives --value-only --search-passes 0"
ncu --set full --clock-control none --import-source on -k regex:search_kernel_serial -s 20 -c 1 -f -o gpurun_out/prof_serial_v1 $CMD > gpurun_out/ncu_serial.log 2>&1
ls -la gpurun_out/prof_serial_v1.ncu-rep
