#!/bin/bash
# round 2, GPU call 1: tests + smoke + bench of the new default build, then the decision-kernel variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_c1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c1.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_c1.log
python bench.py > gpurun_out/bench_c1.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_c1.log
python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/bench_ref_c1.log 2>&1; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref_c1.log
bash profiles/tools/run_variants.sh default f_m4h1 f_m5h1 f_m6h2 f_m4h2 f_m6h0 f_ld64 f_c16 f_c4 f_c2 f_c1 2>&1 | tee gpurun_out/variants_c1.log
