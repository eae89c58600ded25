#!/bin/bash
mkdir -p gpurun_out
for f in gemm norm attention unet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit code $?" >> gpurun_out/test_$f.log
done
timeout 600 python tools/time_unet.py 20 1,4 > gpurun_out/time_unet.log 2>&1
PNP_PROFILE_DUMP=gpurun_out/per_op.json timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
grep -E "passed|failed|parity|rel-L2" gpurun_out/test_*.log; cat gpurun_out/time_unet.log; tail -c 900 gpurun_out/bench.log
