#!/bin/bash
# SELL visit: parity tests, launch-shape lab, bench on both SpMV formats, ncu captures of assembly + SELL kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r3_pytest.log
tail -5 gpurun_out/r3_pytest.log
timeout 300 tools/bin/sell_lab 1000 1000 0 > gpurun_out/r3_lab_quad.txt 2>&1
timeout 300 tools/bin/sell_lab 1000 1000 1 > gpurun_out/r3_lab_tri.txt 2>&1
cat gpurun_out/r3_lab_quad.txt
timeout 600 python bench.py --steps 5 --warmup 3 --tts off > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
tail -3 gpurun_out/r3_bench.err; cat gpurun_out/r3_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather|k_spmv_sell|k_sell_fill|k_sell_detect' -c 8 \
    -f -o gpurun_out/r3_prof python bench.py --steps 1 --warmup 3 --iters 3 --tts off --no-cpu > gpurun_out/r3_ncu.log 2>&1
tail -3 gpurun_out/r3_ncu.log
