"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of GPU time)."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = OrderedDict()
for r in data:
    if len(r) > vi:
        agg.setdefault(r[ki].split('(')[0][-60:], []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
print(f'{"launches":>8s} {"avg us":>9s} {"share":>7s}  kernel')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f'{len(v):8d} {sum(v) / len(v) / 1000:9.1f} {100 * sum(v) / tot:6.1f}%  {k}')
print(f'total {tot / 1000:.1f} us over {sum(len(v) for v in agg.values())} launches')
