#!/bin/bash
mkdir -p gpurun_out
python tools/diag_pnp_rows.py > gpurun_out/r2s8_diag_pnp.log 2>&1; cat gpurun_out/r2s8_diag_pnp.log | tail -12
python tools/time_unet.py 10 16,32 > gpurun_out/r2s8_time.log 2>&1; tail -2 gpurun_out/r2s8_time.log
