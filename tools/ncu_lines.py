"""Instruction counts / stall samples per source line from an `ncu --page source --csv` dump + nvdisasm -g listing."""
import csv, re, sys
from collections import Counter
src_csv, sass = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
line = None; amap = {}
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', l)
    if m and line: amap[int(m.group(1), 16)] = line
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = min(int(r[ix['Address']], 16) for r in data)
c = Counter(); smp = Counter()
for r in data:
    a = int(r[ix['Address']], 16) - base
    n = int(r[ix['Instructions Executed']] or 0)
    ln = amap.get(a, ('?', 0))
    c[ln] += n; smp[ln] += int(r[ix['# Samples']] or 0)
tot = sum(c.values())
print('total instructions', tot)
for ln, n in c.most_common(top): print(str(ln).ljust(34), n, f"{100*n/tot:.1f}%", smp[ln])
