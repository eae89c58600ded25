#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit code $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 1500 gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.log 2>&1; tail -c 600 gpurun_out/bench_reference.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_unet.py 2 4 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_r1c.csv
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 60 -o gpurun_out/gemm_r1c -f python tools/profile_unet.py 1 4 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
