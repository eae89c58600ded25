#!/bin/bash
# round 2, GPU session 24: the pair-mode (tcgen05.mma.cta_group::2) 4096-token self-attention and the FMA-pipe exponentials:
# parity tests, timing of every launch variant, then the UNet-level tests with the pair variant forced
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_attention.py -q -s --timeout 200 -k "tcgen05" > gpurun_out/r2s24_attn_tests.log 2>&1
grep -n "rel-L2\|passed\|failed\|Error\|DEAD\|error" gpurun_out/r2s24_attn_tests.log | cut -c1-260 | tail -30
timeout 200 python tools/run_attn_once.py sweep > gpurun_out/r2s24_attn_sweep.log 2>&1; cut -c1-220 gpurun_out/r2s24_attn_sweep.log | tail -16
if grep -q " passed" gpurun_out/r2s24_attn_tests.log && ! grep -q "failed" gpurun_out/r2s24_attn_tests.log; then
  for poly in 0 3; do
    PNP_ATTN_CLUSTER=3 PNP_ATTN_POLY=$poly timeout 300 python -m pytest tests/test_gpu_unet.py -q -s --timeout 250 > gpurun_out/r2s24_unet_pair_poly$poly.log 2>&1
    grep -n "rel\|passed\|failed\|Error" gpurun_out/r2s24_unet_pair_poly$poly.log | cut -c1-200 | tail -8
  done
fi
