timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 300 python tools/trace_gemm.py > gpurun_out/trace4.log 2>&1; cat gpurun_out/trace4.log
timeout 300 python tools/bench_ops.py > gpurun_out/ops6.log 2>&1; cat gpurun_out/ops6.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
python -c "
import json
d=json.load(open('gpurun_out/bench6.json')); print(d['value'], d['ms_per_step'], d['roofline']['gemm_ms'], d['roofline']['attention_ms'], d['roofline']['glue_ms'], d['clocks'])
"
