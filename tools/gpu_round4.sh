#!/bin/bash
# re-entry visit: parity tests on the restored tree, default bench line, launch list of the same command
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r4_pytest.log 2>&1
tail -6 gpurun_out/r4_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
tail -3 gpurun_out/r4_bench.err; cat gpurun_out/r4_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r4_launches.csv \
    python bench.py --steps 1 --warmup 3 --iters 50 --tts off --no-cpu > gpurun_out/r4_ncu_bench.log 2>&1
tail -2 gpurun_out/r4_ncu_bench.log
