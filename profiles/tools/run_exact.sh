run() { timeout 400 python bench.py --no-cpu-baseline --no-primitives --value-only "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', '->', d['value'], d['kernel_ms_per_step'], d['gpu_launches'])"; }
run --steps 32 --warmup 4
run --steps 32 --warmup 4 --contexts 0
run --steps 32 --warmup 4 --search-passes 0
run --steps 32 --warmup 4 --search-passes 0 --contexts 0
run --steps 64 --warmup 8 --search-passes 0 --contexts 0 --streams 8
run --steps 64 --warmup 8 --search-passes 0 --contexts 0 --streams 12
