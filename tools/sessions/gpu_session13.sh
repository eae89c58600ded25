#!/bin/bash
mkdir -p gpurun_out
for f in gemm norm; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit code $?" >> gpurun_out/test_$f.log
done
timeout 600 python tools/time_unet.py 20 1,4 > gpurun_out/time_unet_pair.log 2>&1
PNP_GEMM_CLUSTER=0 timeout 600 python tools/time_unet.py 20 1,4 > gpurun_out/time_unet_single.log 2>&1
timeout 300 python tools/run_attn_once.py > gpurun_out/attn_once.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:self_attn_tc -c 1 -o gpurun_out/attn_tc_r1 -f python tools/run_attn_once.py > gpurun_out/ncu_attn.log 2>&1
PNP_PROFILE_DUMP=gpurun_out/per_op.json timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/test_*.log | cut -c1-300; tail -n 3 gpurun_out/time_unet_*.log gpurun_out/attn_once.log; tail -c 600 gpurun_out/bench.log
