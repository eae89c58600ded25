#!/bin/bash
# round 2, GPU session 6: re-run of the tests whose tolerances / semantics changed after session 5
mkdir -p gpurun_out
python -m pytest tests/test_gpu_batched.py tests/test_gpu_edict.py tests/test_gpu_pipeline.py tests/test_gpu_cli.py tests/test_gpu_pnp_features.py -q --timeout 1800 -s > gpurun_out/r2s6_pytest.log 2>&1
grep -n "parity\|drift\|passed\|failed\|FAILED\|Error\|pnp features" gpurun_out/r2s6_pytest.log | cut -c1-400 | tail -30
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2s6_bench.log 2>&1; tail -1 gpurun_out/r2s6_bench.log | cut -c1-300; grep -o '"minimal_450": {[^}]*}' gpurun_out/r2s6_bench.log
