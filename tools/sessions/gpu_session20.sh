#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_norm.py tests/test_gpu_unet.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
PNP_PROFILE_DUMP=gpurun_out/per_op.json PNP_GEMM_AUTOTUNE_LOG=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -c 600 gpurun_out/bench.log
