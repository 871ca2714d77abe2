#!/bin/bash
# closing visit: full GPU suite + default bench line
set -u
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r17_pytest.log 2>&1
tail -5 gpurun_out/r17_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
tail -3 gpurun_out/r17_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r17_bench.json"))
t = d["metrics"]["time_to_solution"]["multilevel"]
print("value", d["value"], "e2e", d["e2e"]["value"], "asm_ms", d["metrics"]["assemble_ms"], "tts", t["seconds"], t["iterations"], t["ml_setup_ms"], "roof", d["roofline"]["frac"], d["clocks"])
PY
