#!/bin/bash
# round 2, GPU session 11: three-buffer accumulator rotation for the 128x320 tiles - GEMM tests, UNet parity, timing, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -x --timeout 300 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py -q -x --timeout 600 2>&1 | tail -4
timeout 600 python tools/time_unet.py 10 4,8,16,32 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 --no-image-path --no-config1 > gpurun_out/r2s11_bench.log 2>&1; tail -1 gpurun_out/r2s11_bench.log | cut -c1-400
