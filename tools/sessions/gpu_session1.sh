#!/bin/bash
# First GPU session: environment probe + per-kernel parity tests, each file in its own process.
mkdir -p gpurun_out
{
  nvidia-smi
  echo "cpus: $(nproc)"; lscpu | grep -E "Model name|Socket|Core|Thread" 
  free -g | head -2
  ls -d /root/reference baseline/_ref 2>&1
  python -c "import diffusers" 2>&1 | tail -1
  python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count)"
} > gpurun_out/probe.txt 2>&1
for f in gemm norm attention; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 240 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit code $?" >> gpurun_out/test_$f.log
done
tail -n 30 gpurun_out/test_*.log
