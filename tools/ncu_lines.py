#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
warp instructions executed, stall samples and the top stall reasons.  Usage: ncu_lines.py both.csv [topN]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    hdr = None
    out = []
    for r in rows:
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        if r[0] == "":
            continue  # SASS row
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst = int(d["Instructions Executed"])
            samples = int(d["# Samples"])
            thr = int(d["Thread Instructions Executed"])
        except (KeyError, ValueError):
            continue
        stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
        topstall = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        out.append((inst, samples, thr, r[0], r[1].strip()[:110], topstall))
    tot_i = sum(o[0] for o in out)
    tot_s = sum(o[1] for o in out)
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    print("--- by instructions")
    for o in sorted(out, key=lambda o: -o[0])[:top]:
        print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  L%-4s %s  %s" % (100 * o[0] / tot_i, 100 * o[1] / max(1, tot_s), o[2] / max(1, o[0]), o[3], o[4], o[5]))
    print("--- by stall samples")
    for o in sorted(out, key=lambda o: -o[1])[:top // 2]:
        print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  L%-4s %s  %s" % (100 * o[0] / tot_i, 100 * o[1] / max(1, tot_s), o[2] / max(1, o[0]), o[3], o[4], o[5]))


if __name__ == "__main__":
    main()
