#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel shares.

usage: python tools/summarize_launches.py gpurun_out/<x>_launches.csv profiles/launches_<x>_summary.csv "header comment"
"""
import csv
import re
import sys
from collections import defaultdict


def main(src, dst, note=""):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, float(r["Metric Value"]) / 1e3))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    with open(dst, "w") as f:
        if note:
            for l in note.split("\\n"):
                f.write("# " + l + "\n")
        f.write("# total %.2f ms over %d launches\n" % (tot / 1e3, len(rows)))
        f.write("kernel,launches,total_us,share\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%s,%d,%.1f,%.4f\n' % (n.replace(",", ";"), c, t, t / tot))


if __name__ == "__main__":
    main(*sys.argv[1:4])
