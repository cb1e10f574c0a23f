#!/bin/bash
# usage: run_variants_np.sh NAME...  -- like run_variants.sh but for timing-only builds whose results are wrong on purpose (--no-parity)
for v in "$@"; do
  if [ "$v" = default ]; then unset X264CU_LIB; else export X264CU_LIB=$PWD/x264-i386pic_b200/build/var/lib_$v.so; fi
  python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-primitives --value-only --no-parity --contexts 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['kernel_ms_per_step'])"
done
