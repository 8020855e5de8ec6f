timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log; grep -E "rel-L2|passed|failed|rc=" gpurun_out/tests.log | tail -12
timeout 300 python tools/bench_ops.py > gpurun_out/ops11.log 2>&1; cat gpurun_out/ops11.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err
python -c "
import json
d=json.load(open('gpurun_out/bench12.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['gemm_ms'], d['roofline']['attention_ms'], d['roofline']['glue_ms'], d['clocks'], d['config']['p50_latency_ms'])
"
ECHO_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/bench12_prof.err; head -24 gpurun_out/bench12_prof.err
