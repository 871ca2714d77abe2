#!/bin/bash
# 8-GPU visit: distributed parity at world 8, weak-scaling bench in both comm modes
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r5_ngpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tests/dist_worker.py > gpurun_out/r5_dist8.log 2>&1
grep -E "dist ok|Error|error|assert" gpurun_out/r5_dist8.log | head -20
for mode in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 \
      bench.py --gpus 8 --steps 3 --warmup 3 --tts off --comm $mode > gpurun_out/r5_bench_$mode.json 2> gpurun_out/r5_bench_$mode.err
  tail -3 gpurun_out/r5_bench_$mode.err | cut -c1-300; cat gpurun_out/r5_bench_$mode.json | cut -c1-300
done
