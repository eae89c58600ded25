#!/bin/bash
# round 2, GPU session 2: batch-size dependence diagnostic, VAE tests, the fixed batched tests
mkdir -p gpurun_out
python tools/diag_batch.py > gpurun_out/r2s2_diag.log 2>&1; cat gpurun_out/r2s2_diag.log | tail -9
PNP_GEMM_AUTOTUNE=0 python tools/diag_batch.py > gpurun_out/r2s2_diag_noautotune.log 2>&1; tail -8 gpurun_out/r2s2_diag_noautotune.log
python -m pytest tests/test_gpu_vae.py tests/test_gpu_batched.py -q --timeout 900 -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/r2s2_pytest.log
grep -n "rel-L2\|vs oracle\|vs reference\|passed\|failed\|FAILED\|Error" gpurun_out/r2s2_pytest.log | head -40
