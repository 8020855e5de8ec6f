"""Summarise an ncu --set full report (run here, no GPU): per launch duration, DRAM traffic, tensor-pipe activity,
registers, plus the top stall reasons of the source view. Usage: python tools/ncu_summary.py <report.ncu-rep> [names...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
names = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__cycles_elapsed.avg",
        "launch__grid_size", "launch__cluster_dim_x"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
units = rows[1]
print("# " + rep)
print("# columns: " + ", ".join(f"{w} [{units[i]}]" for w, i in idx))
for n, r in enumerate(rows[2:]):
    label = names[n] if n < len(names) else r[hdr.index("Kernel Name")][:60]
    print(f"{label:34s} " + "  ".join(f"{r[i][:11]:>11s}" for _, i in idx))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for n, s in enumerate(secs):
    h = s["rows"][0]
    data = [r for r in s["rows"][1:] if len(r) == len(h)]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[ix[k]]) for r in data) for k in stalls}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:6]
    label = s["name"].replace("void echo::", "").replace("echo::<unnamed>::", "")[:58]
    if (label, tot) in seen:
        continue
    seen.add((label, tot))
    print(f"## {label}: {tot} samples, {len(data)} SASS instructions; top stalls: " + ", ".join(f"{k[6:]} {v}" for k, v in top))
