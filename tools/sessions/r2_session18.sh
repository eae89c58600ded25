#!/bin/bash
# round 2, GPU session 18: A/B of the cross-attention tiles-per-CTA on kernel durations (ncu, cross_attn launches only)
mkdir -p gpurun_out
for t in 1 2 4 8; do
  PNP_CROSS_TPC=$t timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:cross_attn --csv --log-file gpurun_out/r2_cross_tpc$t.csv python tools/profile_unet.py 1 32 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_cross_tpc$t.csv') if not l.startswith('=='))]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value'); g=h.index('Grid Size')
from collections import defaultdict
d=defaultdict(list)
for r in rows[1:]: d[(r[k].split('<')[-1][:4], r[g])].append(float(r[v].replace(',',''))/1e3)
print('tpc=$t', {kk: (len(vv), round(sum(vv)/len(vv),1)) for kk,vv in d.items()}, 'total us', round(sum(sum(vv) for vv in d.values()),1))
PY
done
