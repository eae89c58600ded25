#!/bin/bash
mkdir -p gpurun_out
PNP_ATTN_CLUSTER=1 timeout 300 python tools/run_attn_once.py 2>&1 | tail -1
timeout 300 python tools/run_attn_once.py 2>&1 | tail -1
PNP_ATTN_CLUSTER=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:self_attn_tc -c 1 -o gpurun_out/attn_tc_r2 -f python tools/run_attn_once.py > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
