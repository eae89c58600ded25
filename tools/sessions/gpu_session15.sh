#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_gemm2.py > gpurun_out/prof_gemm2.log 2>&1
grep "gemm prof\|====" gpurun_out/prof_gemm2.log | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -s --timeout 600 -p no:cacheprovider > gpurun_out/test_attention.log 2>&1
echo "exit code $?" >> gpurun_out/test_attention.log
grep -E "passed|failed|rel-L2|Error|error" gpurun_out/test_attention.log | cut -c1-300
timeout 300 python tools/run_attn_once.py 2>&1 | tail -2
PNP_GEMM_CLUSTER=0 timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
