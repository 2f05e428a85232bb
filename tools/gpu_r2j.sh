#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not config0_z100 and not multirank" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -8 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_c2.json 2> gpurun_out/r2j_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_c2.json'))
print(d['ms_per_step'], d['stage_ms_last_step'])
for k,v in d['stages'].items(): print(k, round(v['ms_per_step'],3), v.get('frac_of_hbm_peak'), v.get('frac_of_fp32_peak'))
PY
