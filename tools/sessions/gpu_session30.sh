#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit code $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
/usr/bin/time -v -o gpurun_out/bench_time.txt timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 2500 gpurun_out/bench_default.log; grep Elapsed gpurun_out/bench_time.txt
/usr/bin/time -v -o gpurun_out/ref_time.txt timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.log 2>&1; tail -c 900 gpurun_out/bench_reference.log; grep Elapsed gpurun_out/ref_time.txt
