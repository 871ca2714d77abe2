#!/bin/bash
# 2-GPU visit: distributed worker (parity incl. resultants + owned rows), 2-rank bench with the owned-rows e2e
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "owned or resultants" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 tests/dist_worker.py > gpurun_out/r14_dist2.log 2>&1
grep -E "dist ok|Error|error|assert" gpurun_out/r14_dist2.log | head -20
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r14_bench_n2.json 2> gpurun_out/r14_bench_n2.err
tail -3 gpurun_out/r14_bench_n2.err | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/r14_bench_n2.json"):
    if l.startswith("{"):
        d = json.loads(l); print("N=2 value", d["value"], "e2e", d["e2e"], "asm", d["metrics"]["assemble_ms"], "tts", d["metrics"]["time_to_solution"]["multilevel"]["seconds"], d["metrics"]["time_to_solution"]["multilevel"]["iterations"])
PY
