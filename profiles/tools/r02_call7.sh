#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_c7.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_c7.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_c7.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_c7.log
