#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-primitives"
$B --workload 64streams > gpurun_out/bench_64streams_1gpu.log 2>&1; echo "64streams rc=$?"
$B --workload 4k > gpurun_out/bench_4k_1gpu.log 2>&1; echo "4k rc=$?"
$B --search-passes 0 > gpurun_out/bench_exact_1gpu.log 2>&1; echo "exact rc=$?"
$B --e2e-arrays eager > gpurun_out/bench_eager_1gpu.log 2>&1; echo "eager rc=$?"
$B --workload 4k-one-stream > gpurun_out/bench_one4k_1gpu.log 2>&1; echo "one4k rc=$?"
$B --workload 1080p-one-stream > gpurun_out/bench_one1080p_1gpu.log 2>&1; echo "one1080p rc=$?"
for f in 64streams 4k exact eager one4k one1080p; do python - <<PY
import json
for line in open("gpurun_out/bench_${f}_1gpu.log"):
    if line.startswith("{"):
        d=json.loads(line); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("d2h_bytes_per_step"), d.get("parity_check",{}) and d["parity_check"].get("ok"))
PY
done
python profiles/tools/x264_cli_bench.py 96 16 medium > gpurun_out/x264_cli_bench_1080p.log 2>&1; cat gpurun_out/x264_cli_bench_1080p.log
python profiles/tools/x264_cli_bench.py 96 16 veryfast > gpurun_out/x264_cli_bench_1080p_veryfast.log 2>&1; cat gpurun_out/x264_cli_bench_1080p_veryfast.log
