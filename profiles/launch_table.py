#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the launch count, total and
share of GPU time (cold-cache, serialised: shares matter, not absolutes).  usage: launch_table.py launches.csv [skip_regex]"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, skip=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
        if skip and re.search(skip, name):
            continue
        a = agg.setdefault(name, [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | share | first grid x block |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % | {a[2]} x {a[3]} |")
    print(f"| total | {sum(a[0] for a in agg.values())} | {tot:.3f} | | |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
