timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k attention > gpurun_out/tests_attn.log 2>&1; echo "rc=$?" >> gpurun_out/tests_attn.log
tail -30 gpurun_out/tests_attn.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
tail -8 gpurun_out/tests.log
timeout 300 python tools/bench_ops.py > gpurun_out/ops4.log 2>&1; tail -3 gpurun_out/ops4.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
python -c "
import json
d=json.load(open('gpurun_out/bench5.json')); print(d['value'], d['ms_per_step'], d['roofline']['gemm_ms'], d['roofline']['attention_ms'], d['roofline']['glue_ms'])
"
