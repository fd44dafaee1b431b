#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one row per kernel (mean over its launches) as markdown, and a
JSON with the measured DRAM traffic per launch that bench.py reports as roofline.traffic."""
import collections
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "ms"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("smsp__inst_executed.sum", "warp inst"), ("dram__bytes_read.sum", "dram rd GB"), ("dram__bytes_write.sum", "dram wr GB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %")]


def to_gb(v, unit):
    v = float(v)
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit, 1.0)


def main(path, md_out, json_out):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
        d = acc.setdefault(name, collections.defaultdict(list))
        for c, _ in COLS:
            if c in h:
                i = h.index(c)
                v = to_gb(r[i], u[i]) if "bytes" in c else float(r[i])
                d[c].append(v)
    lines = ["| kernel | launches | " + " | ".join(t for _, t in COLS) + " |", "|---|---:|" + "---:|" * len(COLS)]
    traffic = {}
    for name, d in acc.items():
        n = len(d[COLS[0][0]])
        cells = []
        for c, _ in COLS:
            m = sum(d[c]) / max(1, len(d[c]))
            cells.append("%.3e" % m if "inst" in c else "%.3f" % m)
        lines.append("| `%s` | %d | " % (name, n) + " | ".join(cells) + " |")
        traffic[name] = {"launches": n, "ms": sum(d["gpu__time_duration.sum"]) / n,
                         "dram_gb_per_launch": (sum(d["dram__bytes_read.sum"]) + sum(d["dram__bytes_write.sum"])) / n}
    open(md_out, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(json_out, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
