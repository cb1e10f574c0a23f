#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_c2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_c2.log
bash profiles/tools/run_variants.sh default g_c1ns g_c1ns_jc g_c1ns_jc_h0m6 g_jc 2>&1 | tee gpurun_out/variants_c2.log
bash profiles/tools/capture.sh r02a 2>&1 | tail -12
