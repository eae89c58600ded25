#!/bin/bash
# round 2, GPU session 4: knob experiments at the batch sizes the bench now uses, ncu launch list + full capture (warm
# caches) of one B=32 forward, the complete -m gpu suite
mkdir -p gpurun_out
echo "== default" > gpurun_out/r2s4_knobs.log; python tools/time_unet.py 10 8,32 >> gpurun_out/r2s4_knobs.log 2>&1
echo "== PNP_GN_CLUSTER=16" >> gpurun_out/r2s4_knobs.log; PNP_GN_CLUSTER=16 python tools/time_unet.py 10 8,32 >> gpurun_out/r2s4_knobs.log 2>&1
echo "== PNP_GN_CLUSTER=0" >> gpurun_out/r2s4_knobs.log; PNP_GN_CLUSTER=0 python tools/time_unet.py 10 8,32 >> gpurun_out/r2s4_knobs.log 2>&1
echo "== PNP_ATTN_CLUSTER=2" >> gpurun_out/r2s4_knobs.log; PNP_ATTN_CLUSTER=2 python tools/time_unet.py 10 8,32 >> gpurun_out/r2s4_knobs.log 2>&1
echo "== PNP_PDL=1" >> gpurun_out/r2s4_knobs.log; PNP_PDL=1 python tools/time_unet.py 10 8,32 >> gpurun_out/r2s4_knobs.log 2>&1
grep -v "model ready" gpurun_out/r2s4_knobs.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b32.csv python tools/profile_unet.py 1 32 > gpurun_out/r2s4_ncu1.log 2>&1
tail -2 gpurun_out/r2s4_ncu1.log; wc -l gpurun_out/r2_launches_b32.csv
timeout 1500 ncu --profile-from-start off --set full --clock-control none --cache-control none -k regex:"gemm_tcgen05|self_attn|gn_|ln_kernel|cross_attn" -c 120 -o /tmp/r2_full_b32 -f python tools/profile_unet.py 1 32 > gpurun_out/r2s4_ncu2.log 2>&1
tail -2 gpurun_out/r2s4_ncu2.log
ncu -i /tmp/r2_full_b32.ncu-rep --page raw --csv > gpurun_out/r2_full_b32_raw.csv 2>/dev/null; ls -la /tmp/r2_full_b32.ncu-rep; wc -c gpurun_out/r2_full_b32_raw.csv
python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -15 > gpurun_out/r2s4_pytest.log; tail -8 gpurun_out/r2s4_pytest.log
