#!/bin/bash
# round 2, GPU session 9: complete -m gpu suite + default bench line on the final build
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1800 -s > gpurun_out/r2s9_pytest_full.log 2>&1
grep -n "split-operand\|50-step parity\|passed\|failed\|FAILED\|Error" gpurun_out/r2s9_pytest_full.log | cut -c1-400 | tail -20
python bench.py --steps 3 --warmup 3 > gpurun_out/r2s9_bench_default.log 2>&1; tail -1 gpurun_out/r2s9_bench_default.log | cut -c1-400
