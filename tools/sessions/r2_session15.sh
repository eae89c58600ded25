#!/bin/bash
# round 2, GPU session 15: launch list with DRAM traffic of one B=32 forward after the raster / GroupNorm changes
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2b_launches_b32.csv python tools/profile_unet.py 1 32 > gpurun_out/r2s15_ncu1.log 2>&1
tail -2 gpurun_out/r2s15_ncu1.log; wc -l gpurun_out/r2b_launches_b32.csv
