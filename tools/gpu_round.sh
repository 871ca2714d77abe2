#!/bin/bash
# One GPU-box visit, parametrised:   gpurun --timeout T [--gpus N] -- 'bash tools/gpu_round.sh TAG STEP [STEP ...]'
# Outputs land in gpurun_out/TAG_*.  Steps:
#   tests        -m gpu suite minus fullsize                     fullsize   tests/test_gpu_fullsize.py -s (prints the residuals)
#   parity       tests/test_gpu_parity.py only           bench      default bench line (N=1)
#   benchq       bench without tts / cpu baseline        ref        reference arm (CPU oracle) at N=1
#   launches     ncu launch list of a short bench        ncu:REGEX[:SKIP[:COUNT]]  ncu --set full of matching kernels
#   asm          tools/asm_time.py (values pass, Tri-3 + Quad-4)
#   dist:N       tests/dist_worker.py on N ranks         benchN:N   bench.py --gpus N under torchrun
#   refN:N       reference arm under torchrun, N ranks   py:FILE    python FILE (a tools/ probe)
set -u
TAG=$1; shift
mkdir -p gpurun_out
O=gpurun_out/$TAG
PORT=29700
NPROF=0
for STEP in "$@"; do
  ARG=${STEP#*:}; KIND=${STEP%%:*}
  echo "=== $STEP"
  case $KIND in
    tests)    ( time timeout 1500 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_fullsize.py ) > ${O}_pytest.log 2>&1; tail -6 ${O}_pytest.log ;;
    fullsize) ( time timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s ) > ${O}_fullsize.log 2>&1; grep -E "passed|failed|error|iterations|block-scaled|Jacobi|assert|Error" ${O}_fullsize.log | cut -c1-400 | tail -40 ;;
    parity)   ( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multilevel.py -m gpu -q -x ) > ${O}_parity.log 2>&1; tail -5 ${O}_parity.log ;;
    bench)    timeout 900 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err | cut -c1-400; python tools/bench_digest.py ${O}_bench.json ;;
    benchq)   timeout 600 python bench.py --steps 5 --warmup 3 --tts off --no-cpu > ${O}_benchq.json 2> ${O}_benchq.err; tail -3 ${O}_benchq.err | cut -c1-400; python tools/bench_digest.py ${O}_benchq.json ;;
    ref)      timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > ${O}_ref.json 2> ${O}_ref.err; tail -3 ${O}_ref.err | cut -c1-400; python tools/bench_digest.py ${O}_ref.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
                  python bench.py --steps 1 --warmup 3 --iters 20 --tts off --no-cpu > ${O}_ncu_bench.log 2>&1; tail -2 ${O}_ncu_bench.log | cut -c1-300 ;;
    ncu)      # ncu:REGEX[:SKIP[:COUNT]] -> ${O}_prof_<n>.ncu-rep
              IFS=: read -r RX SKIP CNT <<< "$ARG"; NPROF=$((NPROF+1))
              timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s ${SKIP:-0} -c ${CNT:-3} -f -o ${O}_prof_$NPROF \
                  python bench.py --steps 1 --warmup 3 --iters 5 --tts off --no-cpu > ${O}_ncu_full_$NPROF.log 2>&1; grep -E "PROF|Error" ${O}_ncu_full_$NPROF.log | tail -3 | cut -c1-200 ;;
    asm)      timeout 300 python tools/asm_time.py 2>&1 | tail -6 | tee ${O}_asm.txt ;;
    dist)     PORT=$((PORT+1)); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ARG --master-addr 127.0.0.1 --master-port $PORT tests/dist_worker.py > ${O}_dist$ARG.log 2>&1
              grep -E "dist ok|Error|error|assert" ${O}_dist$ARG.log | head -20 | cut -c1-400 ;;
    benchN)   PORT=$((PORT+1)); timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ARG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $ARG --steps 5 --warmup 3 > ${O}_bench_n$ARG.json 2> ${O}_bench_n$ARG.err
              tail -3 ${O}_bench_n$ARG.err | cut -c1-400; python tools/bench_digest.py ${O}_bench_n$ARG.json ;;
    refN)     PORT=$((PORT+1)); timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ARG --master-addr 127.0.0.1 --master-port $PORT bench.py --impl reference --gpus $ARG --steps 3 --warmup 1 > ${O}_ref_n$ARG.json 2> ${O}_ref_n$ARG.err
              tail -3 ${O}_ref_n$ARG.err | cut -c1-400; python tools/bench_digest.py ${O}_ref_n$ARG.json ;;
    py)       timeout 900 python $ARG 2>&1 | tail -40 | cut -c1-400 | tee ${O}_$(basename $ARG .py).txt ;;
    *)        echo "unknown step $STEP" ;;
  esac
done
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -8
ls -la gpurun_out | tail -30
