"""Print the top stall-sampled SASS instructions of the first kernel in an `ncu --page source --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
blk = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
print(blk[0][1])
hdr = blk[1]
si, ci, ei = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
data = [r for r in blk[2:] if len(r) > si]
tot = sum(float(r[si]) for r in data)
print('total samples', tot, 'instructions', len(data))
top = sorted(range(len(data)), key=lambda i: -float(data[i][si]))[:n]
for i in sorted(top):
    r = data[i]
    print(f'{i:5d} {100 * float(r[si]) / tot:6.2f}% exec={r[ei]:>8s}  {r[ci].strip()[:110]}')
