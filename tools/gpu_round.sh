#!/bin/bash
# one GPU-box visit: parity tests, bench, ncu launch list, ncu full captures of the hot kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --tts-max-s 30 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
# launch list of a short bench run (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --iters 20 --tts off --no-cpu > gpurun_out/ncu_bench.log 2>&1
# full capture of the SpMV and the coloured assembly kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spmv|k_assemble_gather|k_assemble_colored|k_update|k_direction' -s 30 -c 12 \
    -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --iters 5 --tts off --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
