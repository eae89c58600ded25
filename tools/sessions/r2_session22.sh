#!/bin/bash
# round 2, GPU session 22 (after the container was re-created): complete -m gpu suite with per-test durations (the driver's
# limit for this step is 1200 s) + the default bench line on the rebuilt library
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1800 --durations=30 > gpurun_out/r2s22_pytest_full.log 2>&1
grep -n "passed\|failed\|FAILED\|Error\|skipped" gpurun_out/r2s22_pytest_full.log | cut -c1-300 | tail -12
grep -n "slowest" -A32 gpurun_out/r2s22_pytest_full.log | cut -c1-160 | head -40
python bench.py --steps 3 --warmup 3 > gpurun_out/r2s22_bench_default.log 2>&1; tail -1 gpurun_out/r2s22_bench_default.log | cut -c1-400
