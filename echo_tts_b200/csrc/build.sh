#!/bin/bash
# Builds libecho_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -I../../include ${ECHO_NVCC_EXTRA}"
BUILD=${ECHO_BUILD_DIR:-build}
OUT=${ECHO_OUT:-../libecho_b200.so}
mkdir -p $BUILD
pids=()
for f in *.cu; do
  o=$BUILD/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ] || [ ../../include/echo_b200.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT $BUILD/*.o -lcudart
echo "built $OUT"
