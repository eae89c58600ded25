#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_gemm.py > gpurun_out/prof_gemm.log 2>&1
grep "gemm prof\|====" gpurun_out/prof_gemm.log | awk 'NR%2==0 || /====/' | cut -c1-400
