#!/bin/bash
# round 2, GPU session 25: warp-role layout with the MMA / TMA roles on the highest warp ids (issue priority) and remote
# arrivals without the cluster-scope release fence (pair mode): parity of every tcgen05 attention variant under both role
# layouts, variant timing at B = 32 and 4, UNet-level tests with the new layout (default cluster mode and pair mode)
mkdir -p gpurun_out
for r in 0 1; do
  PNP_ATTN_ROLES=$r timeout 200 python -m pytest tests/test_gpu_attention.py -q --timeout 150 -k "tcgen05" > gpurun_out/r2s25_attn_tests_roles$r.log 2>&1
  tail -2 gpurun_out/r2s25_attn_tests_roles$r.log | cut -c1-200
done
timeout 200 python tools/run_attn_once.py sweep 32 4 > gpurun_out/r2s25_attn_sweep.log 2>&1; grep "attn prof" gpurun_out/r2s25_attn_sweep.log | awk 'NR%2==0' | cut -c1-230
grep "variant" gpurun_out/r2s25_attn_sweep.log | cut -c1-120
PNP_ATTN_ROLES=1 timeout 200 python -m pytest tests/test_gpu_unet.py -q --timeout 150 > gpurun_out/r2s25_unet_roles1.log 2>&1; tail -1 gpurun_out/r2s25_unet_roles1.log
PNP_ATTN_ROLES=1 PNP_ATTN_CLUSTER=3 timeout 200 python -m pytest tests/test_gpu_unet.py -q --timeout 150 > gpurun_out/r2s25_unet_roles1_pair.log 2>&1; tail -1 gpurun_out/r2s25_unet_roles1_pair.log
