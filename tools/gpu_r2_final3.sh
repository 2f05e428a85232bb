#!/bin/bash
# round-2 final state, one GPU: the full -m gpu suite and the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -10 gpurun_out/r2f_pytest.log
timeout 600 python bench.py > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_c2.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'] and d['cpu_baseline']['ms_per_step'], d.get('halofind_peaks'))
for k,v in d['stages'].items(): print(k, round(v['ms_per_step'],3), v.get('frac_of_hbm_peak'), v.get('frac_of_fp32_peak'))
PY
