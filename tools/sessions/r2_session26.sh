#!/bin/bash
# round 2, GPU session 26: event-driven MMA issue order in the 4096-token self-attention (PNP_ATTN_SCHED=1): parity of every
# tcgen05 attention variant, variant timing at B = 32 and 4, UNet-level tests (default cluster mode and pair mode)
mkdir -p gpurun_out
PNP_ATTN_SCHED=1 timeout 200 python -m pytest tests/test_gpu_attention.py -q --timeout 150 -k "tcgen05" > gpurun_out/r2s26_attn_tests_sched1.log 2>&1
tail -2 gpurun_out/r2s26_attn_tests_sched1.log | cut -c1-200
timeout 200 python tools/run_attn_once.py sweep 32 4 > gpurun_out/r2s26_attn_sweep.log 2>&1; grep "attn prof" gpurun_out/r2s26_attn_sweep.log | awk 'NR%2==0' | cut -c1-200
grep "variant" gpurun_out/r2s26_attn_sweep.log | cut -c1-125
PNP_ATTN_SCHED=1 timeout 200 python -m pytest tests/test_gpu_unet.py -q --timeout 150 > gpurun_out/r2s26_unet_sched1.log 2>&1; tail -1 gpurun_out/r2s26_unet_sched1.log
PNP_ATTN_SCHED=1 PNP_ATTN_CLUSTER=3 timeout 200 python -m pytest tests/test_gpu_unet.py -q --timeout 150 > gpurun_out/r2s26_unet_sched1_pair.log 2>&1; tail -1 gpurun_out/r2s26_unet_sched1_pair.log
