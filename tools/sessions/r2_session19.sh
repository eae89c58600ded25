#!/bin/bash
# round 2, GPU session 19: occupancy of the statistics and LayerNorm kernels (registers), ncu durations of the norm kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_norm.py tests/test_gpu_unet.py -q -x --timeout 600 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"gn_|ln_kernel|vt_transpose|cross_attn" --csv --log-file gpurun_out/r2_norm_launches.csv python tools/profile_unet.py 1 32 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_norm_launches.csv') if not l.startswith('=='))]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[1:]: d[r[k].split('::')[-1].split('(')[0]].append(float(r[v].replace(',',''))/1e3)
for kk,vv in sorted(d.items()): print(kk, len(vv), 'total %.1f us' % sum(vv), 'max %.1f' % max(vv))
PY
