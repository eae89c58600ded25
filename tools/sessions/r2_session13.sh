#!/bin/bash
# round 2, GPU session 13: N-fastest tile order for GEMMs whose activations exceed L2
mkdir -p gpurun_out
PNP_GEMM_RASTER=1 timeout 600 python -m pytest tests/test_gpu_gemm.py -q -x --timeout 300 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py -q -x --timeout 300 2>&1 | tail -3
for mode in 0 1 -1; do
  echo "PNP_GEMM_RASTER=$mode"
  PNP_GEMM_RASTER=$mode timeout 600 python tools/time_unet.py 10 8,16,32 2>&1 | grep "B="
done
