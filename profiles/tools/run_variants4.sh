#!/bin/bash
# usage: run_variants4.sh NAME...   ("default" = the in-tree library); per tuning build: per-kernel ms on ONE context and the
# frames/s of the benchmarked configuration (4 contexts whose kernels overlap, so launch tails are filled)
for v in "$@"; do
  if [ "$v" = default ]; then unset X264CU_LIB; else export X264CU_LIB=$PWD/x264-i386pic_b200/build/var/lib_$v.so; fi
  python bench.py --steps 48 --warmup 8 --no-cpu-baseline --no-primitives --value-only 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value4', d['value'], d['kernel_ms_per_step'])"
done
