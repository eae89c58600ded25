#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -s --timeout 300 -p no:cacheprovider 2>&1 | grep -E "passed|failed|rel-L2|rror" | cut -c1-300
PNP_ATTN_PROF=1 timeout 300 python tools/run_attn_once.py 2>&1 | grep "attn prof" | head -2 | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_masactrl.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
