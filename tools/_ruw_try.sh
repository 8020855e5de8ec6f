set -u
for rep in 1 2; do
for lib in echo_tts_b200/libecho_b200_base.so echo_tts_b200/libecho_b200.so; do
echo "== $lib"
ECHO_B200_LIB=$PWD/$lib timeout 600 python tools/bench_ops.py --only attention 2>&1 | grep -E "attention"
done
done
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -2
