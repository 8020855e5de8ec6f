set -u
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k fused_residual 2>&1 | tail -4
ECHO_DAC_RU_WINDOW=0 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k fused_residual 2>&1 | tail -2
echo "== window form"; timeout 300 python tools/trace_dac_ru.py 2>&1 | tail -30
echo "== bench"; timeout 300 python tools/bench_dac_ru.py 2>&1 | tail -8
echo "== ring"; ECHO_DAC_RU_WINDOW=0 timeout 300 python tools/bench_dac_ru.py 2>&1 | head -3
