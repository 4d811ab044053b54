#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total, mean, share.
Usage: ncu_launches.py launches.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[d["Metric Unit"]]
            a = agg[d["Kernel Name"][:70]]
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s n=%4d total %10.1f us  mean %9.1f us  share %5.1f%%" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


if __name__ == "__main__":
    main()
