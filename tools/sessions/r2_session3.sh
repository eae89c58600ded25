#!/bin/bash
# round 2, GPU session 3: where do batch sizes diverge (layer diagnostic), early 50-step parity, new batch tests,
# bench sweep at larger image batches + the masactrl / edict workloads
mkdir -p gpurun_out
python tools/diag_layers.py > gpurun_out/r2s3_diag_layers.log 2>&1; tail -30 gpurun_out/r2s3_diag_layers.log
python tools/parity50_partial.py > gpurun_out/r2s3_parity50.log 2>&1; tail -3 gpurun_out/r2s3_parity50.log
python -m pytest tests/test_gpu_masactrl.py tests/test_gpu_edict.py tests/test_gpu_gemm.py tests/test_gpu_norm.py -q --timeout 900 -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2s3_pytest.log
grep -n "rel-L2\|batched vs\|passed\|failed\|FAILED\|Error" gpurun_out/r2s3_pytest.log | head -30
: > gpurun_out/r2s3_bench.log
for cfg in "p2p 4 1" "p2p 6 1" "p2p 8 1" "p2p 8 2" "masactrl 4 1" "edict 8 1"; do
  set -- $cfg
  echo "== workload $1 batch $2 lanes $3" >> gpurun_out/r2s3_bench.log
  timeout 900 python bench.py --workload $1 --batch $2 --lanes $3 --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/r2s3_bench.log 2>&1
done
grep -o '"value": [0-9.]*\|== workload.*\|"e2e": {"value": [0-9.]*\|Error.*' gpurun_out/r2s3_bench.log
