#!/bin/bash
# round 2, GPU session 14: traversal direction of the non-GEMM passes (PNP_REVERSE bits) and the raster threshold
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 600 python tools/time_unet.py 10 8,32 2>&1 | grep "B="; }
run PNP_REVERSE=0
run PNP_REVERSE=1
run PNP_REVERSE=3
run PNP_REVERSE=15
run PNP_REVERSE=31 PNP_GN_APPLY=1
run PNP_REVERSE=16 PNP_GN_APPLY=1
run PNP_REVERSE=0 PNP_GEMM_RASTER_MB=8
PNP_REVERSE=15 timeout 600 python -m pytest tests/test_gpu_norm.py tests/test_gpu_attention.py tests/test_gpu_unet.py -q -x --timeout 300 2>&1 | tail -3
