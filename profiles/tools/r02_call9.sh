#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_mc_surfaces.py -m gpu -q -x > gpurun_out/pytest_c9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_c9.log
python profiles/tools/time_prim.py r02
bash profiles/tools/capture_prims.sh r02b 2>&1 | tail -8
