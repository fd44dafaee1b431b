#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        name = re.sub(r"^\(anonymous namespace\)::", "", name)
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val if unit in ("ms", "msecond") else val * 1e3
        rows.append((name, ms, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ms, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values()) or 1.0
    print("| kernel | launches | total ms | avg ms | share | grid (first) | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.4f | %.1f %% | %s | %s |" % (name, a[0], a[1], a[1] / a[0], 100 * a[1] / total, a[2], a[3]))
    print("\ntotal %.3f ms over %d launches" % (total, len(rows)))


if __name__ == "__main__":
    main(sys.argv[1])
