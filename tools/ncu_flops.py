#!/usr/bin/env python
"""Measured FP32 FLOPs of one frame's trace kernels from an ncu CSV log taken with
    --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum
(SURVEY 8d source (1): FLOPs = fadd + fmul + 2 ffma).  Sums every kernel between two cull_kernel launches except the cull and
resolve kernels themselves, prints the per-kernel split and merges {workload: flops} into profiles/flops.json.
usage: ncu_flops.py <log.csv> <workload> [frame_index]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ii, ki, mi, vi = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    d = launches.setdefault(int(r[ii]), {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(',', ''))
ls = list(launches.values())
cuts = [i for i, l in enumerate(ls) if 'cull_kernel' in l["name"]] + [len(ls)]
f = int(sys.argv[3]) if len(sys.argv) > 3 else max(0, len(cuts) - 3)  # default: the last complete frame
frame = ls[cuts[f]:cuts[f + 1]]
tot, per = 0.0, collections.OrderedDict()
for l in frame:
    if 'cull_kernel' in l["name"] or 'resolve_kernel' in l["name"]:
        continue
    fl = sum(v * (2.0 if 'ffma' in k else 1.0) for k, v in l.items() if k.startswith('smsp__sass_thread_inst_executed_op_f'))
    short = re.sub(r'\(.*', '', l["name"]).replace('void ', '')
    short = re.sub(r'<.*', '', short) + (('<' + re.search(r'<\(?(?:gort::PoolSrc\))?(\d)', l["name"]).group(1) + '>') if 'pool_trace' in l["name"] else '')
    per[short] = per.get(short, 0.0) + fl
    tot += fl
for k, v in per.items():
    print("%-34s %.4g FLOP" % (k, v))
print("frame %d: %d launches, measured FP32 FLOPs (fadd + fmul + 2 ffma) = %.6g" % (f, len(frame), tot))
path = os.path.join(ROOT, "profiles", "flops.json")
try:
    data = json.load(open(path))
except (OSError, ValueError):
    data = {}
data[sys.argv[2]] = tot
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
