#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/w_train_launches.csv python scripts/train_profile.py --ncu > gpurun_out/w_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/w_train_launches.csv')))
hdr=None; d=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        try:
            k=r[hdr.index('Kernel Name')].split('(')[0][:55]; v=float(r[hdr.index('Metric Value')]); d.setdefault(k,[]).append(v)
        except: pass
tot=sum(sum(v) for v in d.values())
print('total %.2f ms, launches %d'%(tot/1e6, sum(len(v) for v in d.values())))
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1]))[:8]:
    print(f'{k:55s} n={len(v):4d} total={sum(v)/1e6:8.3f} ms {100*sum(v)/tot:5.1f}%')
PY
