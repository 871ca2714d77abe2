#!/bin/bash
# 8-GPU visit: north-star target (4000x4000-node Tri-3 plate, 96 M DOF) with the multilevel preconditioner,
# weak-scaling bench with multilevel time-to-solution, distributed parity at world 8
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
free -g | head -2
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 tools/target_run.py --nodes 4000 --kind t > gpurun_out/r9_target_c3.json 2> gpurun_out/r9_target_c3.err
echo "target rc=$?"; tail -5 gpurun_out/r9_target_c3.err | cut -c1-300; cat gpurun_out/r9_target_c3.json | cut -c1-2500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --tts-pc ml > gpurun_out/r9_bench_n8.json 2> gpurun_out/r9_bench_n8.err
echo "bench rc=$?"; tail -3 gpurun_out/r9_bench_n8.err | cut -c1-300; cat gpurun_out/r9_bench_n8.json | cut -c1-2500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29643 tests/dist_worker.py > gpurun_out/r9_dist8.log 2>&1
grep -E "dist ok|Error|error|assert" gpurun_out/r9_dist8.log | head -20
