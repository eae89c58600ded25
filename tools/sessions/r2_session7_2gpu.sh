#!/bin/bash
# round 2, GPU session 7 (2 GPUs): the torchrun paths - bench.py as the driver launches it, and the sharded CLI sweep
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2s7_bench2.log 2>&1
tail -1 gpurun_out/r2s7_bench2.log | cut -c1-600
python -c "from pnpinversion_b200 import cli; cli.write_synthetic_dataset('/tmp/pie', n_items=5, size=512)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 run_editing_p2p.py --data_path /tmp/pie --output_path /tmp/pie_out --num_ddim_steps 3 --batch 2 --edit_method_list directinversion+p2p > gpurun_out/r2s7_cli2.log 2>&1
grep -n '"rank"' gpurun_out/r2s7_cli2.log; find /tmp/pie_out -name "*.jpg" | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --no-config1 > gpurun_out/r2s7_ref2.log 2>&1
tail -1 gpurun_out/r2s7_ref2.log | cut -c1-200
python -m pytest tests/test_gpu_pnp_features.py tests/test_gpu_batched.py -q --timeout 900 -s > gpurun_out/r2s7_pytest.log 2>&1; grep -n "pnp rows|passed|failed|FAILED" gpurun_out/r2s7_pytest.log | cut -c1-300
