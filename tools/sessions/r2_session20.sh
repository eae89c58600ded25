#!/bin/bash
# round 2, GPU session 20: cta_group::2 pair tiles as autotune candidates at B=32 (PNP_GEMM_CLUSTER=1) vs the default, by
# the ncu durations of the GEMM launches of one forward
mkdir -p gpurun_out
for c in 0 1; do
  PNP_GEMM_CLUSTER=$c timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/r2_gemm_cluster$c.csv python tools/profile_unet.py 1 32 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_gemm_cluster$c.csv') if not l.startswith('=='))]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[1:]: d[r[k].split('::')[-1].split('(')[0]].append(float(r[v].replace(',',''))/1e3)
print('PNP_GEMM_CLUSTER=$c', {kk: (len(vv), round(sum(vv)/1e3,3)) for kk,vv in sorted(d.items())}, 'total ms', round(sum(sum(vv) for vv in d.values())/1e3,3))
PY
done
