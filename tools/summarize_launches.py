"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid) count / total / share."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"^void |echo::|\(anonymous namespace\)::", "", name)
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
    key = (name, r[ix["Grid Size"]])
    agg[key][0] += 1
    agg[key][1] += us
    total += us
print(f"total {total/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us/1e3:9.3f} ms {100*us/total:5.1f}%  n={n:5d} avg={us/n:8.1f} us  grid={grid:>14s}  {name[:90]}")
