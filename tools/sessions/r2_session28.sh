#!/bin/bash
# round 2, GPU session 28 (the last 80 GPU-seconds): cost of a tcgen05.commit in the MMA issue stream
mkdir -p gpurun_out
timeout 60 python tools/mma_commit_probe.py > gpurun_out/r2_mma_commit_probe.txt 2>&1; cat gpurun_out/r2_mma_commit_probe.txt | cut -c1-160
