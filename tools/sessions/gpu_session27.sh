#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
for L in 1 2 3; do
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes $L > gpurun_out/bench_l$L.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_l$L.log') if l.startswith('{')][-1])
    print('lanes $L: value %.4f e2e %.4f img/s ms/step %.1f clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks']))
except Exception as e:
    print('lanes $L failed', e); print(open('gpurun_out/bench_l$L.log').read()[-1500:])
PY
done
