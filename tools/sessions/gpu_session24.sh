#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py tests/test_gpu_pipeline.py -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/time_unet.py 20 1,4 2>&1 | tail -2
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_unet.py 2 4 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_r1c.csv
timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tcgen05 -c 40 -o /tmp/gemm_r1c -f python tools/profile_unet.py 1 4 > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
ncu -i /tmp/gemm_r1c.ncu-rep --page raw --csv > gpurun_out/gemm_r1c_raw.csv 2>/dev/null
ls -la /tmp/gemm_r1c.ncu-rep gpurun_out/
PNP_PROFILE_DUMP=gpurun_out/per_op.json timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -c 400 gpurun_out/bench.log
du -sh gpurun_out
