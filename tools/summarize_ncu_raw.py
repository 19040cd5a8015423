#!/usr/bin/env python
"""Summarise an `ncu -i X.ncu-rep --page raw --csv` export of the convolution launches of one step.

usage: tools/summarize_ncu_raw.py raw.csv [layer-names...]  -> markdown table on stdout
Columns: kernel, device time, DRAM bytes read / written (dram__bytes_*.sum), issue-slot utilisation.
"""
import csv
import re
import sys

LAYERS = (["encoder.conv1"] + [f"encoder.dense1.layers.{i}" for i in range(4)] + ["encoder.dense1.transition",
          "encoder.conv2"] + [f"encoder.dense2.layers.{i}" for i in range(4)] + ["encoder.dense2.transition",
          "encoder.conv3"] + [f"encoder.dense3.layers.{i}" for i in range(4)] +
          ["encoder.dense3.transition (pass 0)", "encoder.dense3.transition (pass 1)", "encoder.conv4", "decoder.conv1",
           "decoder.conv2", "decoder.conv3", "decoder.conv4"] + [f"decoder.final_dense.layers.{i}" for i in range(4)] +
          ["decoder.final_dense.transition"])

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {c: i for i, c in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    print("| # | layer | kernel | time us | DRAM read MB | DRAM write MB | issue active % |\n|---|---|---|---|---|---|---|")
    tot_t = tot_b = dense_t = dense_b = 0.0
    for k, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "").strip()
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        ia = float(r[col["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]])
        layer = LAYERS[k] if len(data) == len(LAYERS) else ""
        tot_t += t
        tot_b += rd + wr
        if ".layers." in layer:
            dense_t += t
            dense_b += rd + wr
        print(f"| {k} | {layer} | `{name}` | {t:.0f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | {ia:.0f} |")
    print(f"\nAll {len(data)} launches: {tot_t:.0f} us, {tot_b / 1e9:.2f} GB DRAM traffic ({int(tot_b)} bytes).")
    if dense_t:
        print(f"The 16 dense-block 3x3 launches: {dense_t:.0f} us, {dense_b / 1e9:.2f} GB DRAM traffic ({int(dense_b)} bytes).")


if __name__ == "__main__":
    main()
