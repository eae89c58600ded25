#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -3
PNP_GEMM_AUTOTUNE_LOG=1 timeout 600 python tools/time_unet.py 20 1,4 > gpurun_out/time_unet_tuned.log 2>&1
grep "gemm tune" gpurun_out/time_unet_tuned.log | sort | uniq -c | sort -k3,3 -k4,4 | cut -c1-200
tail -n 2 gpurun_out/time_unet_tuned.log
PNP_GEMM_AUTOTUNE=0 timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
PNP_ATTN_CLUSTER=1 PNP_PROFILE_DUMP=gpurun_out/per_op.json timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -c 900 gpurun_out/bench.log
