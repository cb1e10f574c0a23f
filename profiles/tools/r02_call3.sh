#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bench_config.py tests/test_slicetype_decisions.py tests/test_gpu_lookahead.py -m gpu -q -x > gpurun_out/pytest_c3.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_c3.log
bash profiles/tools/run_variants.sh default h_m6 h_m4h1 h_m5h1 h_c4 h_c16 h_c1 2>&1 | tee gpurun_out/variants_c3.log
