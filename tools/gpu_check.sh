#!/bin/bash
# One gpurun call: GPU parity tests, a short bench run, (optionally) the eager-PyTorch context number.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tests|bench|oracle|all] [pytest args]'
set -u
what=${1:-all}
shift || true
mkdir -p gpurun_out
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 1500 python -m pytest tests -m gpu -q -s "$@" > gpurun_out/tests.log 2>&1
  echo "pytest rc=$?" | tee -a gpurun_out/tests.log
  tail -25 gpurun_out/tests.log
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"
  tail -c 6000 gpurun_out/bench.json
  tail -5 gpurun_out/bench.err
fi
if [ "$what" = oracle ] || [ "$what" = all ]; then
  timeout 600 python tools/bench_oracle_cuda.py --runs 3 > gpurun_out/oracle_cuda.txt 2>&1
  tail -6 gpurun_out/oracle_cuda.txt
fi
