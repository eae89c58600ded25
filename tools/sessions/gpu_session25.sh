#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k geglu --timeout 600 -p no:cacheprovider 2>&1 | tail -2
PNP_PROFILE_DUMP=gpurun_out/per_op.json timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -c 300 gpurun_out/bench.log
timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
timeout 600 python tools/time_unet.py 20 4 2>&1 | tail -1
