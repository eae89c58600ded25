#!/bin/bash
# round 2, GPU session 5: the complete -m gpu suite with every fixture (50-step pipeline, masactrl pipeline, pnp features,
# CLIs), the default bench line as the driver runs it, the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1800 -s > gpurun_out/r2s5_pytest_full.log 2>&1
grep -n "parity\|rel-L2\|vs reference\|vs oracle\|passed\|failed\|FAILED\|Error\|minimal-350" gpurun_out/r2s5_pytest_full.log | cut -c1-400 | tail -60
python bench.py --steps 3 --warmup 3 > gpurun_out/r2s5_bench_default.log 2>&1; tail -1 gpurun_out/r2s5_bench_default.log | cut -c1-1500
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s5_bench_reference.log 2>&1; tail -1 gpurun_out/r2s5_bench_reference.log | cut -c1-900
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s5_smoke.log 2>&1; tail -2 gpurun_out/r2s5_smoke.log
