python -m pytest tests -m gpu -x -q -k "exact or serial or cli or raster" 2>&1 | tail -3
for v in default ser4; do
  if [ "$v" = default ]; then unset X264CU_LIB; else export X264CU_LIB=$PWD/x264-i386pic_b200/build/var/lib_$v.so; fi
  python bench.py --no-cpu-baseline --no-primitives --search-passes 0 --value-only 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v exact4', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['parity_check']['ok'])"
  python bench.py --no-cpu-baseline --no-primitives --search-passes 0 --streams 8 --value-only 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v exact8', d['value'], d['ms_per_step'])"
done
