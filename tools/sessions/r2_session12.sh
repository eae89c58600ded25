#!/bin/bash
# round 2, GPU session 12: vectorised GroupNorm apply pass (channel vector per thread, fast SiLU, back-to-front traversal)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_norm.py tests/test_gpu_unet.py tests/test_gpu_vae.py -q -x --timeout 600 2>&1 | tail -4
for mode in 0 1 3; do
  echo "PNP_GN_APPLY=$mode"
  PNP_GN_APPLY=$mode timeout 600 python tools/time_unet.py 10 8,16,32 2>&1 | grep "B="
done
