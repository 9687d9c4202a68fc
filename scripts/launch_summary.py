"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list (second forward only)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != 'ID']
names = [r[4].split('(')[0] for r in rows]
# keep the last forward: from the last embed_kernel on
start = max(i for i, n in enumerate(names) if n.endswith('embed_kernel'))
d = collections.OrderedDict()
for r, n in list(zip(rows, names))[start:]:
    d.setdefault(n, []).append(float(r[-1]) / 1e3)
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print(f'{k:40s} n={len(v):3d} avg {sum(v) / len(v):8.1f} us  total {sum(v):8.1f}  {sum(v) / tot * 100:5.1f}%')
print(f'launches {sum(len(v) for v in d.values())}  total {tot:.1f} us')
