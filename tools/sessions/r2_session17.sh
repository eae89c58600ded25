#!/bin/bash
# round 2, GPU session 17: streaming cross-attention kernel (K/V staged once per CTA, query tiles double-buffered)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_attention.py tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_batched.py -q -x --timeout 900 2>&1 | tail -4
echo "== new"; timeout 600 python tools/time_unet.py 10 4,8,32 2>&1 | grep "B="
