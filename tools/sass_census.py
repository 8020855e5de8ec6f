"""Per-kernel SASS opcode census of libecho_b200.so (cuobjdump -sass): the Blackwell-native evidence -- tcgen05 MMA
(UTCHMMA / UTCQMMA), TMA loads (UTMALDG), TMEM loads / stores (LDTM / STTM), tcgen05 commit barriers (UTCBAR), and the
fire-and-forget fp32 reductions (REDG), and the absence of legacy mma.sync (HMMA / IMMA) anywhere. usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "echo_tts_b200", "libecho_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern = None
counts = collections.OrderedDict()
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU.EX2", "REDG", "HMMA", "IMMA"]
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"echo::\(anonymous namespace\)::|echo::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[kern]["_total"] += 1
    for k in KEYS:
        if (op.startswith(k) if k != "HMMA" else op.startswith("HMMA")):
            counts[kern][k] += 1
print(f"# SASS census of echo_tts_b200/libecho_b200.so ({os.path.getsize(lib)} bytes), cuobjdump -sass, sm_100a")
print(f"# columns: instructions | " + " ".join(KEYS))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[x] for x in KEYS):
        print(f"{c['_total']:7d} | " + " ".join(f"{x}={c[x]}" for x in KEYS if c[x]) + f" | {k}")
print("# kernels without any of the above opcodes (plain CUDA-core glue): " + str(sum(1 for c in counts.values() if not any(c[x] for x in KEYS))))
print("# totals: " + " ".join(f"{x}={tot[x]}" for x in KEYS))
print(f"# legacy tensor-core opcodes (HMMA / IMMA = mma.sync): {tot['HMMA'] + tot['IMMA']}")
