set -u
timeout 900 python -m pytest tests/test_dac_gpu.py tests/test_dac_encode_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/phase_times.py 2>&1 | tail -12
