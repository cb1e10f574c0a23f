#!/bin/bash
# final 1-GPU evidence of round 2: GPU tests, smoke, the default bench line, ncu launch list + full captures of both search passes and
# the decision kernel, then the other single-GPU workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref_final.log
profiles/tools/capture.sh r02e 2>&1 | tail -4
profiles/tools/r02_call11.sh 2>&1 | tail -20
