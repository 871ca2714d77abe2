#!/bin/bash
# round-1 h: full GPU suite, default bench line, launch list of the same command, ncu --set full of the assembly + SpMV + vector kernels
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r13_pytest.log 2>&1
tail -5 gpurun_out/r13_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r13_bench.json 2> gpurun_out/r13_bench.err
tail -3 gpurun_out/r13_bench.err; cut -c1-400 gpurun_out/r13_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r13_launches.csv \
    python bench.py --steps 1 --warmup 3 --iters 50 --tts off --no-cpu > gpurun_out/r13_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather|k_spmv_sell|k_update|k_direction' -s 12 -c 6 \
    -f -o gpurun_out/r13_full python bench.py --steps 1 --warmup 3 --iters 3 --tts off --no-cpu > gpurun_out/r13_ncu_full.log 2>&1
tail -2 gpurun_out/r13_ncu_full.log | cut -c1-200
python tools/asm_time.py 2>&1 | tail -2
