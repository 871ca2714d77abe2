#!/bin/bash
# assembly kernel iteration: parity subset + bench (assembly numbers), optional tri timing
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r12_pytest.log 2>&1
tail -4 gpurun_out/r12_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --tts off --no-cpu > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err
tail -3 gpurun_out/r12_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r12_bench.json"))
print("assemble_ms", d["metrics"]["assemble_ms"], "el/s", d["metrics"]["elements_assembled_per_s"], "value", d["value"], "e2e", d["e2e"]["value"])
PY
timeout 300 python tools/asm_time.py 2>&1 | tail -4
