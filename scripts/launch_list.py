"""Prints (kernel, grid, block, ns) of an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi, gi, bi, mi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size'), H.index('Block Size'), H.index('Metric Name')
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = [(r[ki][:70], r[gi], r[bi], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi and r[mi] == 'gpu__time_duration.sum']
for o in out[-tail:]:
    print('%-72s %-14s %-12s %9.0f' % o)
