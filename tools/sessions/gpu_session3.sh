#!/bin/bash
mkdir -p gpurun_out
for f in epilogue unet pipeline; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit code $?" >> gpurun_out/test_$f.log
done
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
# one full-metric capture of the three heaviest kernels (B=4 forward): conv GEMM, self-attention d=40, GroupNorm apply
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|self_attn_kernel|gn_apply|ln_kernel" -s 900 -c 40 -o gpurun_out/prof_r1a python tools/time_unet.py 1 4 > gpurun_out/ncu_full.log 2>&1
tail -n 12 gpurun_out/test_*.log gpurun_out/bench.log gpurun_out/smoke.log
