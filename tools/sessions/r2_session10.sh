#!/bin/bash
# round 2, GPU session 10: default bench line incl. the image-path end-to-end figure; quick re-run of the editor / VAE tests
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r2s10_bench_default.log 2>&1; tail -1 gpurun_out/r2s10_bench_default.log | cut -c1-300
grep -o '"image_path_e2e": {[^}]*}' gpurun_out/r2s10_bench_default.log; grep -o '"minimal_450": {"value": [0-9.]*' gpurun_out/r2s10_bench_default.log
tail -3 gpurun_out/r2s10_bench_default.log | grep -i "error\|Traceback" 
python -m pytest tests/test_gpu_vae.py tests/test_gpu_cli.py -q --timeout 900 2>&1 | tail -3
