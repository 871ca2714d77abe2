#!/bin/bash
# 2-GPU box visit: distributed parity test + weak-scaling bench at N=2 (and N=1 for the ratio)
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15 > gpurun_out/n2_pytest.log
tail -5 gpurun_out/n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 5 --warmup 3 --tts off > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -3 gpurun_out/n2_bench.err; cat gpurun_out/n2_bench.json
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --tts off --no-cpu > gpurun_out/n2_bench1.json 2>> gpurun_out/n2_bench.err
cat gpurun_out/n2_bench1.json
