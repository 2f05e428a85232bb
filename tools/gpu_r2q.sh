#!/bin/bash
# PP_EXT tests + c2 bench (stage table)
mkdir -p gpurun_out
T=${1:-r2q}
timeout 900 python -m pytest tests -m gpu -x -q -k "pp_ext or clustered or pair_force or smoke or replay" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_c2.json'))
print(d['ms_per_step'], d['stage_ms_last_step'])
for k,v in d['stages'].items(): print(k, round(v['ms_per_step'],3), v.get('frac_of_hbm_peak'), v.get('frac_of_fp32_peak'))
PY
