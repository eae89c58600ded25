#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_masactrl.py -m gpu -x -q -s --timeout 600 -p no:cacheprovider 2>&1 | grep -E "passed|failed|rel-L2|rror" | cut -c1-200
timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench.py default took ${SECONDS}s"; tail -c 3000 gpurun_out/bench_default.log
SECONDS=0
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.log 2>&1; echo "reference arm took ${SECONDS}s"; tail -c 1200 gpurun_out/bench_reference.log
