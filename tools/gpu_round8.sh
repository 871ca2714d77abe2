#!/bin/bash
# 2-GPU visit: weak-scaling bench with multilevel time-to-solution
set -u
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --tts-pc ml > gpurun_out/r8_bench_n2.json 2> gpurun_out/r8_bench_n2.err
echo "rc=$?"
tail -15 gpurun_out/r8_bench_n2.err | cut -c1-400
cut -c1-3000 gpurun_out/r8_bench_n2.json
