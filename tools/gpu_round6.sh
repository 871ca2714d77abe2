#!/bin/bash
# one GPU-box visit (session 4): parity tests, bench with full time-to-solution, ncu full captures with source
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r6_pytest_gpu.log
tail -5 gpurun_out/r6_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --tts-max-s 150 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
tail -3 gpurun_out/r6_bench.err; cat gpurun_out/r6_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv|k_assemble_gather|k_update|k_direction' -s 30 -c 10 \
    -f -o gpurun_out/r6_prof python bench.py --steps 1 --warmup 3 --iters 5 --tts off --no-cpu > gpurun_out/r6_ncu_full.log 2>&1
ls -la gpurun_out
