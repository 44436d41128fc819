"""Aggregate ncu stall samples (source page CSV of a --set full report) per CUDA source line, using nvdisasm -g line info.
usage: python tools/ncu_lines.py src.csv kernel.sass [file-substring]"""
import csv, re, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; si = h.index('# Samples'); src = h.index('Source')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
ncu = [(r[src].strip(), int(r[si] or 0), r) for r in rows[2:] if len(r) > si]
cur = None; sass = []
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: sass.append((m.group(2).strip(), cur))
assert len(sass) == len(ncu), (len(sass), len(ncu))
main = sys.argv[3] if len(sys.argv) > 3 else '.cu'
ctx = None; c = Counter(); cs = {}
for (t, n, r), (s, ln) in zip(ncu, sass):
    if ln and main in ln[0]: ctx = ln[1]
    c[ctx] += n
    d = cs.setdefault(ctx, Counter())
    for j in stall_cols: d[h[j]] += int(r[j] or 0)
tot = sum(c.values())
print('total samples', tot)
for k, v in sorted(c.items(), key=lambda x: -x[1])[:28]:
    print(f"{v:6d} {100*v/tot:5.1f}%  line {k}  {cs[k].most_common(2)}")
