#!/bin/bash
# assembly kernel under ncu --set full (one launch) with source correlation
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather' -s 3 -c 1 \
    -f -o gpurun_out/r11_asm python bench.py --steps 1 --warmup 3 --iters 3 --tts off --no-cpu > gpurun_out/r11_ncu.log 2>&1
tail -3 gpurun_out/r11_ncu.log
ls -la gpurun_out/
