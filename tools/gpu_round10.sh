#!/bin/bash
# resultants visit: full GPU parity suite (incl. the new stress-resultant tests and the CLI twins), 2-rank dist worker if 2 GPUs
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r10_pytest.log 2>&1
tail -8 gpurun_out/r10_pytest.log
