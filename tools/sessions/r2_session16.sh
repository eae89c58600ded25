#!/bin/bash
# round 2, GPU session 16: register-resident GroupNorm for low-resolution levels of image batches, 3-deep residual prefetch in
# the GEMM epilogue, vectorised V transposition
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_norm.py tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_unet.py -q -x --timeout 600 2>&1 | tail -4
echo "== new"; timeout 600 python tools/time_unet.py 10 4,8,32 2>&1 | grep "B="
echo "== PNP_GN_LOCAL=0"; PNP_GN_LOCAL=0 timeout 600 python tools/time_unet.py 10 8,32 2>&1 | grep "B="
echo "== new again"; timeout 600 python tools/time_unet.py 10 8,32 2>&1 | grep "B="
