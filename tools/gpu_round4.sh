#!/bin/bash
# 2-GPU visit: peer-window comm vs NCCL (parity worker + bench in both modes)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r4_pytest.log
tail -12 gpurun_out/r4_pytest.log
for mode in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
      bench.py --gpus 2 --steps 5 --warmup 3 --tts off --comm $mode > gpurun_out/r4_bench_$mode.json 2> gpurun_out/r4_bench_$mode.err
  tail -3 gpurun_out/r4_bench_$mode.err | cut -c1-300; cat gpurun_out/r4_bench_$mode.json | cut -c1-400
done
