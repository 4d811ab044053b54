"""Per-kernel totals of one frame from an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python tools/launch_breakdown.py launches.csv [frame_index]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
recs = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hi + 1:] if len(r) > vi and r[vi]]
idx = [i for i, (k, _) in enumerate(recs) if 'cull_kernel' in k] + [len(recs)]
f = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fr = recs[idx[f]:idx[f + 1]]


def name(k):
    k2 = re.sub(r'\(.*', '', k)
    k2 = re.sub(r'<.*', '', k2).replace('void ', '')
    if 'pool_trace' in k:
        k2 += '<' + re.search(r'<\(?(?:gort::PoolSrc\))?(\d)', k).group(1) + '>'
    return k2


tot = collections.OrderedDict()
for k, v in fr:
    k2 = name(k)
    tot.setdefault(k2, [0, 0.0])
    tot[k2][0] += 1
    tot[k2][1] += v
for k, (n, v) in tot.items():
    print("%-34s n=%4d total %10.1f us" % (k, n, v / 1000))
print('sum ms %.3f over %d launches' % (sum(v for _, v in fr) / 1e6, len(fr)))
