#!/bin/bash
# round 2, GPU session 1: full -m gpu suite (new: C loops, batching, clone, span mapper, batched LocalBlend, config-1
# golden), UNet timing vs batch size, bench sweep over (images per pass, concurrent passes)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/r2s1_pytest.log
tail -5 gpurun_out/r2s1_pytest.log
python tools/time_unet.py 20 1,3,4,8,12,16 > gpurun_out/r2s1_time_unet.log 2>&1
cat gpurun_out/r2s1_time_unet.log | tail -8
: > gpurun_out/r2s1_bench.log
for cfg in "1 1" "1 3" "3 1" "3 2" "4 2"; do
  set -- $cfg
  echo "== batch $1 lanes $2" >> gpurun_out/r2s1_bench.log
  timeout 600 python bench.py --batch $1 --lanes $2 --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/r2s1_bench.log 2>&1
done
grep -o '"value": [0-9.]*\|== batch.*\|"e2e": {"value": [0-9.]*' gpurun_out/r2s1_bench.log
