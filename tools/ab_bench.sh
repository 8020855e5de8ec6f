#!/bin/bash
# A/B of two builds of libecho_b200.so on the same box: tools/ab_bench.sh <lib_a> <lib_b> [bench args]
# (boxes differ by +-4 %, so a change is only judged against a control run in the same gpurun call)
A=$1; B=$2; shift 2
for rep in 1 2; do
  for lib in "$A" "$B"; do
    ECHO_B200_LIB=$lib timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'audio-s/s', d['clocks']['sm_mhz'])"
  done
done
