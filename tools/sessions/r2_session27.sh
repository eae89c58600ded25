#!/bin/bash
# round 2, GPU session 27: event-driven MMA issue order with NON-BLOCKING barrier probes (mbarrier.test_wait; session 26
# polled with try_wait, whose failed probes suspend the thread: 9 ms per layer): parity, timing at B = 32, UNet-level tests
mkdir -p gpurun_out
PNP_ATTN_SCHED=1 timeout 120 python -m pytest tests/test_gpu_attention.py -q --timeout 100 -k "tcgen05" > gpurun_out/r2s27_attn_tests_sched1.log 2>&1
tail -1 gpurun_out/r2s27_attn_tests_sched1.log | cut -c1-200
timeout 120 python tools/run_attn_once.py sweep 32 > gpurun_out/r2s27_attn_sweep.log 2>&1; grep "attn prof" gpurun_out/r2s27_attn_sweep.log | awk 'NR%2==0' | cut -c1-420
grep "variant" gpurun_out/r2s27_attn_sweep.log | cut -c1-125
PNP_ATTN_SCHED=1 timeout 120 python -m pytest tests/test_gpu_unet.py -q --timeout 100 > gpurun_out/r2s27_unet_sched1.log 2>&1; tail -1 gpurun_out/r2s27_unet_sched1.log
PNP_ATTN_SCHED=1 PNP_ATTN_CLUSTER=3 timeout 120 python -m pytest tests/test_gpu_unet.py -q --timeout 100 > gpurun_out/r2s27_unet_sched1_pair.log 2>&1; tail -1 gpurun_out/r2s27_unet_sched1_pair.log
