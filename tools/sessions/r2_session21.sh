#!/bin/bash
# round 2, GPU session 21: complete -m gpu suite + default bench line on the build with the session 12-19 kernel changes
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1800 -s > gpurun_out/r2s21_pytest_full.log 2>&1
grep -n "masactrl 50-step\|50-step parity\|passed\|failed\|FAILED\|Error\|skipped" gpurun_out/r2s21_pytest_full.log | cut -c1-400 | tail -20
python bench.py --steps 3 --warmup 3 > gpurun_out/r2s21_bench_default.log 2>&1; tail -1 gpurun_out/r2s21_bench_default.log | cut -c1-300
