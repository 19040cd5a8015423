#!/usr/bin/env python
"""Summarise an `ncu -i X.ncu-rep --page raw --csv` export of ALL launches of one step (tools/r02_capture.sh).

usage: tools/summarize_step_raw.py raw.csv  -> markdown table on stdout
Per launch: device time, DRAM bytes read / written, DRAM throughput, tensor-pipe activity, issue-slot utilisation.
"""
import csv
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {c: i for i, c in enumerate(hdr)}

    def val(r, name, default=float("nan")):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return default
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    tensor = next((c for c in ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_active",
                               "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                               "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active") if c in col), None)
    tensor_cols = [c for c in hdr if "pipe_tensor" in c]
    print(f"tensor-pipe columns present: {tensor_cols}\n")
    print("| # | kernel | grid | time us | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe % | issue active % |\n|---|---|---|---|---|---|---|---|---|")
    tot_t = tot_b = 0.0
    groups = {}
    for k, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("cdan::", "").replace("<unnamed>::", "").strip()
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        ia = val(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed")
        tp = val(r, tensor) if tensor else float("nan")
        tot_t += t
        tot_b += rd + wr
        g = groups.setdefault(re.sub(r"<.*", "", name), [0.0, 0.0, 0])
        g[0] += t; g[1] += rd + wr; g[2] += 1
        print(f"| {k} | `{name}` | {r[col['Grid Size']]} | {t:.0f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | {(rd + wr) / t / 1e3:.0f} | {tp:.0f} | {ia:.0f} |")
    print(f"\nAll {len(data)} launches: {tot_t:.0f} us (serialised, cold cache), {tot_b / 1e9:.2f} GB DRAM traffic.\n")
    print("| kernel family | launches | time us | share | DRAM GB | GB/s |\n|---|---|---|---|---|---|")
    for n, (t, b, c) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        print(f"| `{n}` | {c} | {t:.0f} | {100 * t / tot_t:.1f} % | {b / 1e9:.2f} | {b / t / 1e3:.0f} |")


if __name__ == "__main__":
    main()
