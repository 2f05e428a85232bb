#!/bin/bash
# full single-GPU test suite (without the long configs[0] run) + c2 bench
mkdir -p gpurun_out
T=${1:-r2r}
timeout 1200 python -m pytest tests -m gpu -x -q -k "not config0_z100 and not multirank" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_c2.json'))
print(d['ms_per_step'], d['stage_ms_last_step'])
for k,v in d['stages'].items(): print(k, round(v['ms_per_step'],3), v.get('frac_of_hbm_peak'), v.get('frac_of_fp32_peak'))
PY
