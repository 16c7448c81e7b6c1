#!/usr/bin/env python
"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`) per kernel.
usage: python tools/ncu_launches.py launches.csv [--seq]   (--seq also prints the launch sequence)"""
import csv
import sys
from collections import OrderedDict


def main(path, seq=False):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, order = OrderedDict(), []
    for r in rows[1:]:
        name = r[kn].split("(")[0]
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[mu], 1.0)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
        order.append((name, v))
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,max_us,share_pct")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%s,%d,%.1f,%.1f,%.1f" % (k, a[0], a[1], a[2], 100 * a[1] / tot))
    if seq:
        print("# sequence")
        for k, v in order:
            print("%-40s %9.1f" % (k, v))


if __name__ == "__main__":
    main(sys.argv[1], "--seq" in sys.argv)
