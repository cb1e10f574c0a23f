#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_c8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_c8.log
python bench.py --no-cpu-baseline > gpurun_out/bench_c8.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
for line in open("gpurun_out/bench_c8.log"):
    if line.startswith("{"):
        d=json.loads(line)
        print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_step")}); print(d["e2e"]); print(json.dumps(d["primitives"]))
PY
bash profiles/tools/capture_prims.sh r02 2>&1 | tail -4
