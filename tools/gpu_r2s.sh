#!/bin/bash
mkdir -p gpurun_out
for w in c1 c1c; do
CUBEP3M_B200_SANDWICH=v2 timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/r2s_bench_$w.json 2> gpurun_out/r2s_bench_$w.err; echo "bench $w rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2s_bench_$w.json'))
print(d['ms_per_step'], d['stage_ms_last_step'])
for k,v in d['stages'].items():
    if v['ms_per_step']>0.05: print(k, round(v['ms_per_step'],3), round(v['us_per_launch'],1), v.get('frac_of_hbm_peak'))
PY
done
