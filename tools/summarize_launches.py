#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the captured steps)."""
import collections
import csv
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, tot = collections.OrderedDict(), 0.0
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("cdan::<unnamed>::", "").replace("unnamed>::", "")
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
    tot += v
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {c} | {v / 1e3:.1f} | {100 * v / tot:.1f} % |")
print(f"| **all** | {sum(c for _, c in agg.values())} | {tot / 1e3:.1f} | 100 % |")
