#!/bin/bash
# usage: run_variants_exact.sh NAME...  -- exact raster-serial mode: one context (latency) and one context per stream (throughput)
for v in "$@"; do
  if [ "$v" = default ]; then unset X264CU_LIB; else export X264CU_LIB=$PWD/x264-i386pic_b200/build/var/lib_$v.so; fi
  for c in 1 0; do
  python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-primitives --value-only --search-passes 0 --contexts $c 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v contexts=$c', d['value'], d['kernel_ms_per_step'])"
  done
done
