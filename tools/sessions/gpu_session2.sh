#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/test_unet.log 2>&1
echo "exit code $?" >> gpurun_out/test_unet.log
timeout 600 python tools/time_unet.py 20 1,4 > gpurun_out/time_unet.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_b4.csv python tools/time_unet.py 2 4 > gpurun_out/ncu_launches.log 2>&1
tail -n 40 gpurun_out/test_unet.log gpurun_out/time_unet.log
