#!/bin/bash
# round 2, GPU session 23: (1) tcgen05.mma instruction-cost probe (is the ~100-cycle floor of the d=40 attention MMAs paid
# per instruction or per SM?), (2) the fused CLIP text encoder tests, (3) the bench with the prompts encoded on the GPU +
# the single-image figure, (4) ncu --set full of the attention / GEMM / norm launches of one B=32 forward on the final build
mkdir -p gpurun_out
timeout 300 python tools/mma_probe.py 1024 > gpurun_out/r2_mma_probe.txt 2>&1; cat gpurun_out/r2_mma_probe.txt | cut -c1-150
timeout 600 python -m pytest tests/test_gpu_clip.py -q -s --timeout 500 > gpurun_out/r2s23_clip.log 2>&1
grep -n "rel-L2\|passed\|failed\|Error\|error" gpurun_out/r2s23_clip.log | cut -c1-250 | tail -15
timeout 600 python bench.py --steps 2 --warmup 3 --text-encoder clip --no-cpu-baseline --no-image-path > gpurun_out/r2s23_bench_clip.log 2>&1
tail -1 gpurun_out/r2s23_bench_clip.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'single_image') if k in d}, d['config'].get('text_encoder'))
except Exception as e:
    print('bench line unreadable', e)
"
tail -3 gpurun_out/r2s23_bench_clip.log | cut -c1-300 | head -2
timeout 700 ncu --profile-from-start off --set full --clock-control none --cache-control none -k regex:"gemm_tcgen05|self_attn|gn_|ln_kernel|cross_attn" -c 100 -o /tmp/r2b_full_b32 -f python tools/profile_unet.py 1 32 > gpurun_out/r2s23_ncu.log 2>&1
tail -2 gpurun_out/r2s23_ncu.log | cut -c1-200
ncu -i /tmp/r2b_full_b32.ncu-rep --page raw --csv > gpurun_out/r2b_full_b32_raw.csv 2>/dev/null; wc -c gpurun_out/r2b_full_b32_raw.csv
python tools/ncu_summary.py gpurun_out/r2b_full_b32_raw.csv gpurun_out/r2b_ncu_full_b32_summary.csv "ncu --set full --clock-control none --cache-control none (warm caches), first 100 GEMM / attention / norm launches of one B=32 UNet forward, final round-2 build (tools/sessions/r2_session23.sh)"
