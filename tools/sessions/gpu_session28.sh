#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_all_r1c.csv python tools/profile_unet.py 1 4 > gpurun_out/ncu_gemm_all.log 2>&1
tail -1 gpurun_out/ncu_gemm_all.log; wc -l gpurun_out/gemm_all_r1c.csv
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 4 > gpurun_out/bench_l4.log 2>&1
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_l4.log') if l.startswith('{')][-1])
    print('lanes 4: value %.4f e2e %.4f img/s ms/step %.1f clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks']))
except Exception as e:
    print('lanes 4 failed', e); print(open('gpurun_out/bench_l4.log').read()[-1500:])
PY
