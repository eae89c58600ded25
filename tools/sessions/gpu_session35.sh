#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"gemm_tcgen05|self_attn_tc" -c 36 -o /tmp/final_r1d -f python tools/profile_unet.py 1 4 > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log
ncu -i /tmp/final_r1d.ncu-rep --page raw --csv > gpurun_out/final_r1d_raw.csv 2>/dev/null
ls -la /tmp/final_r1d.ncu-rep; wc -c gpurun_out/final_r1d_raw.csv
