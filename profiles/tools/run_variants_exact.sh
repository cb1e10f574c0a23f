#!/bin/bash
# usage: run_variants_exact.sh NAME...  -- like run_variants.sh for the exact raster-serial mode (one context)
for v in "$@"; do
  if [ "$v" = default ]; then unset X264CU_LIB; else export X264CU_LIB=$PWD/x264-i386pic_b200/build/var/lib_$v.so; fi
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-primitives --value-only --search-passes 0 --contexts 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['kernel_ms_per_step'])"
done
