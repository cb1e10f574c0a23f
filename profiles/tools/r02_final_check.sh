python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['parity_check']['ok'], d['cpu_baseline']['value'])"
python bench.py --no-cpu-baseline --no-primitives --search-passes 0 > gpurun_out/bench_exact_1gpu.log 2>&1; tail -1 gpurun_out/bench_exact_1gpu.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('exact', d['value'], d['e2e']['value'], d['parity_check']['ok'])"
