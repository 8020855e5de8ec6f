"""Hottest SASS instructions of each kernel in an ncu --set full --import-source on report (run here, no GPU):
samples, top stall reasons and the CUDA source line. Usage: python tools/ncu_hot.py <report.ncu-rep> [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for s in secs:
    h = s["rows"][0]
    ix = {k: i for i, k in enumerate(h)}
    data = [r for r in s["rows"][1:] if len(r) == len(h)]
    if "# Samples" not in ix:
        continue
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"## {s['name'][:90]}: {tot} samples, {len(data)} instructions")
    srccol = ix.get("Source", 1)
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:topn]
    for i in sorted(order):
        r = data[i]
        st = sorted(((k[6:], int(r[ix[k]])) for k in stalls if int(r[ix[k]]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"  #{i:5d} {int(r[ix['# Samples']]):5d}  {r[srccol][:70]:70s} {st}")
