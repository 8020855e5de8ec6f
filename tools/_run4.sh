timeout 300 python tools/trace_gemm.py > gpurun_out/trace3.log 2>&1; cat gpurun_out/trace3.log
timeout 300 python tools/bench_ops.py > gpurun_out/ops5.log 2>&1; cat gpurun_out/ops5.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q > gpurun_out/tests_ops.log 2>&1; tail -3 gpurun_out/tests_ops.log
