"""profiles/r01_ncu_traffic.json from `ncu --set full` reports of tools/ncu_gemm.py (run here, no GPU).
Usage: python tools/ncu_traffic.py <m1920.ncu-rep> <m640.ncu-rep> <commit> > profiles/r01_ncu_traffic.json
The capture is `-k regex:"gemm_tc|attn_tc" -s 10 -c 5`: third round of qkvg, w13, wo, w2, attention."""
import csv
import io
import json
import subprocess
import sys

names = ["gemm_qkvg", "gemm_w13", "gemm_wo", "gemm_w2", "attn_tc"]
out = {"source": f"ncu --set full --clock-control none, tools/ncu_gemm.py, commit {sys.argv[3]}; per launch, cold cache",
       "kernels": {}}
for rep, tag in ((sys.argv[1], "M1920"), (sys.argv[2], "M640")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]

    def col(r, k, scale_from_unit=True):
        i = h.index(k)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        if scale_from_unit:
            v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "usecond": 1.0,
                  "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
        return v

    for n, r in zip(names, rows[2:]):
        out["kernels"][f"{n}_{tag}"] = {
            "kernel": r[h.index("Kernel Name")][:80],
            "time_us": col(r, "gpu__time_duration.sum"),
            "dram_bytes_read": col(r, "dram__bytes_read.sum"),
            "dram_bytes_write": col(r, "dram__bytes_write.sum"),
            "tensor_pipe_active_pct": col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
            "registers": int(col(r, "launch__registers_per_thread", False)),
            "grid": int(col(r, "launch__grid_size", False)),
        }
print(json.dumps(out, indent=1))
