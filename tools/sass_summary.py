"""profiles/sass_summary.md: per kernel of libmmgan_b200.so, how often the Blackwell-native SASS mnemonics occur
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk copies, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, HMMA = legacy mma.sync -- must be 0).   python tools/sass_summary.py > profiles/sass_summary.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gan-des-midi-music-gen_b200", "libmmgan_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "USETMAXREG"]
per = OrderedDict()
cur = None
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
        per[cur] = Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        per[cur]["_total"] += 1
        if op in keys:
            per[cur][op] += 1
print("# SASS evidence (cuobjdump -sass of the shipped libmmgan_b200.so, sm_100a)\n")
print("Counts of static instructions per kernel.  `UTCHMMA` = `tcgen05.mma.kind::f16`, `LDTM` / `STTM` = `tcgen05.ld` / `tcgen05.st`, `UTMALDG` / `UTMASTG` = TMA tensor load / store,")
print("`UBLKCP` = `cp.async.bulk`, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier operations; `HMMA` (legacy `mma.sync`) must not appear.\n")
print("| kernel | instructions | " + " | ".join(keys) + " |")
print("|---|---:|" + "---:|" * len(keys))
tot = Counter()
for k, c in per.items():
    if not any(c[x] for x in keys):
        continue
    print(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[x]) for x in keys) + " |")
    tot.update(c)
print(f"| **all kernels with any of these** | {tot['_total']} | " + " | ".join(str(tot[x]) for x in keys) + " |")
print(f"\n{len(per)} kernels in the library; the others are SIMT (rasteriser, fp32 layers, Adam, BCE, packing).")
