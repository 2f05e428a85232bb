#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pp_ext or clustered or pair_force" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
timeout 600 python tools/profile_pp_clustered.py 3 2>&1 | tail -2
timeout 600 python bench.py --workload c1x --steps 5 --warmup 3 --no-cpu --evolve-to-z 2.0 > gpurun_out/r2n_bench_c1x_z2.json 2> gpurun_out/r2n_bench_c1x_z2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench_c1x_z2.json'))
print(d['ms_per_step'], d['stage_ms_last_step'])
for k,v in d['stages'].items():
    if k.startswith('pp'): print(k, round(v['ms_per_step'],3), v.get('frac_of_fp32_peak'), v.get('pairs_per_s'))
PY
