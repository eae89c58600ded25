#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
timeout 600 python tools/time_unet.py 20 4 2>&1 | tail -1
PNP_PROFILE_DUMP=gpurun_out/per_op.json timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])
print('value %.4f e2e %.4f'%(d['value'],d['e2e']['value']), d['roofline']['by_kernel_ms_per_b4_unet_call'], d['roofline']['traffic'], d['clocks'])
PY
