#!/bin/bash
# round-1 i: full GPU suite, default bench line, launch list of the same command, ncu --set full of the kernels that changed
set -u
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r16_pytest.log 2>&1
tail -5 gpurun_out/r16_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err
tail -3 gpurun_out/r16_bench.err; cut -c1-300 gpurun_out/r16_bench.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r16_launches.csv \
    python bench.py --steps 1 --warmup 3 --iters 50 --tts off --no-cpu > gpurun_out/r16_ncu_bench.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather|k_sell_fill_t|k_spmv_sell' -s 4 -c 4 \
    -f -o gpurun_out/r16_full python bench.py --steps 1 --warmup 3 --iters 2 --tts off --no-cpu > gpurun_out/r16_ncu_full.log 2>&1
tail -2 gpurun_out/r16_ncu_full.log | cut -c1-200
