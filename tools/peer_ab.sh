#!/bin/bash
# N-GPU A/B of the CG exchange paths: peer windows vs NCCL, with the library's phase timings
#   gpurun --gpus N -- 'bash tools/peer_ab.sh TAG N'
TAG=$1; N=${2:-2}; O=gpurun_out/$TAG; mkdir -p gpurun_out
run() {  # name, env..., -- extra bench flags
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  PORT=$((29800 + RANDOM % 100))
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 5 --warmup 3 --tts off --no-cpu --c3 off --no-parity "$@" > ${O}_$name.json 2> ${O}_$name.err
  python - <<PY
import json
d=[json.loads(l) for l in open("${O}_$name.json") if l.startswith("{")][-1]
print("$name", "value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), json.dumps(d["details"].get("peer_wait_us_per_iteration_rank0")))
PY
}
run peer FS_TIMING=1 --
grep -E "peer_window|peer window" ${O}_peer.err | head -40
run nccl FS_X=1 -- --comm nccl
