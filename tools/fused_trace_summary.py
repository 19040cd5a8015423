"""Summarise a CDAN_FUSED_TRACE dump (tools/fused_trace.py, trace build of dense_fused.cu): row periods and E0's row chain."""
import sys
lines = [l.split() for l in open(sys.argv[1]) if l.strip() and not l.startswith("FUSED")]
d = {l[0]: [int(v) for v in l[1:]] for l in lines if len(l) > 10}
rows = range(20, 60)
for k in ("L0_issued", "E0_start", "E3_done", "loader"):
    df = [d[k][i + 1] - d[k][i] for i in rows if d[k][i] > 0 and d[k][i + 1] > 0]
    print(f"{k:12s} row period mean {sum(df) / len(df):6.0f}")
seq = ["E0_start", "E0_drained", "E0_accfree", "E0_versions", "E0_tsterm", "E0_done"]
seq = [s for s in seq if any(v > 0 for v in d.get(s, []))]
for a, b in zip(seq, seq[1:]):
    xs = [d[b][i] - d[a][i] for i in rows if d[a][i] > 0 and d[b][i] > 0]
    print(f"{a}->{b}: mean {sum(xs) / len(xs):.0f} min {min(xs)} max {max(xs)}")
for c in range(4):
    xs = [d[f"E{c}_done"][i] - d[f"E{c}_start"][i] for i in rows if d[f"E{c}_start"][i] > 0]
    ys = [d[f"E{c}_start"][i + 1] - d[f"E{c}_done"][i] for i in rows if d[f"E{c}_start"][i + 1] > 0]
    print(f"E{c}: busy {sum(xs) / len(xs):.0f}  idle {sum(ys) / len(ys):.0f}")
