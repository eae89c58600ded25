#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gemm.py tests/test_gpu_unet.py -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
PNP_ATTN_PROF=1 timeout 300 python tools/run_attn_once.py 2>&1 | grep "attn prof" | head -1 | cut -c1-600
timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
timeout 600 python tools/time_unet.py 20 4 2>&1 | tail -1
