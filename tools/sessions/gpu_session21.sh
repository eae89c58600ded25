#!/bin/bash
PNP_ATTN_PROF=1 timeout 300 python tools/run_attn_once.py 2>&1 | grep "attn prof" | head -3 | cut -c1-700
PNP_ATTN_PROF=1 PNP_ATTN_CLUSTER=2 timeout 300 python tools/run_attn_once.py 2>&1 | grep "attn prof" | head -2 | cut -c1-700
