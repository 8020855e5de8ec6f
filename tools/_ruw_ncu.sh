set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k fused_residual 2>&1 | tail -2
timeout 300 python tools/bench_dac_ru.py 2>&1 | tail -8
timeout 300 python tools/trace_dac_ru.py 2>&1 | grep -A7 "dilation 9"
cat > /tmp/ru_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from echo_tts_b200 import ops
def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to("cuda", dtype)
for C, T in ((96, 1310720), (192, 655360)):
    a = rnd((T, C), 1); w7, w1 = rnd((C, 7 * C), 2, (7 * C) ** -0.5), rnd((C, C), 3, C ** -0.5)
    b7, b1 = rnd((C,), 4, 0.1, torch.float32), rnd((C,), 5, 0.1, torch.float32)
    al2, alo = torch.exp(0.3 * rnd((C,), 6, 1, torch.float32)), torch.exp(0.3 * rnd((C,), 7, 1, torch.float32))
    x = rnd((T, C), 8, 1, torch.float32); nxt = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.residual_unit(a, w7, b7, al2, w1, b1, x, alo, nxt, 3)
    torch.cuda.synchronize()
print("ok")
PY
python /tmp/ru_one.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 2 -c 1 -f -o gpurun_out/r02_ruw96 python /tmp/ru_one.py > gpurun_out/ruw_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 5 -c 1 -f -o gpurun_out/r02_ru192 python /tmp/ru_one.py >> gpurun_out/ruw_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
