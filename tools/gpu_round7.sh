#!/bin/bash
# GPU-box visit (session 5): first run of the multilevel preconditioner -- its parity tests, the whole gpu suite, bench with ML time-to-solution
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multilevel.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r7_pytest_ml.log
tail -25 gpurun_out/r7_pytest_ml.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multilevel.py 2>&1 | tail -8 > gpurun_out/r7_pytest_gpu.log
tail -4 gpurun_out/r7_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --tts-pc ml > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
tail -5 gpurun_out/r7_bench.err; cat gpurun_out/r7_bench.json
