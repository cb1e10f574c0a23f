#!/bin/bash
# usage: r02_multi.sh N   (under gpurun --gpus N): stream-sharded bench, one-stream 4K bench, and (N == 2) the sharding parity tests
N=$1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$T bench.py --gpus $N --steps 48 --warmup 8 --no-primitives > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench rc=$?"
$T bench.py --gpus $N --steps 48 --warmup 8 --workload 4k-one-stream > gpurun_out/bench_one4k_${N}gpu.log 2>&1; echo "one4k rc=$?"
$T bench.py --gpus $N --steps 48 --warmup 8 --workload 1080p-one-stream > gpurun_out/bench_one1080p_${N}gpu.log 2>&1; echo "one1080p rc=$?"
for f in bench_${N}gpu bench_one4k_${N}gpu bench_one1080p_${N}gpu; do python - <<PY
import json
for line in open("gpurun_out/$f.log"):
    if line.startswith("{"):
        d=json.loads(line); print("$f", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
if [ "$N" = 2 ] || [ "$N" = 4 ]; then python -m pytest tests/test_gpu_shard2.py -m gpu -q > gpurun_out/pytest_shard2.log 2>&1; echo "shard tests rc=$?"; tail -3 gpurun_out/pytest_shard2.log; fi
