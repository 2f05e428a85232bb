#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pp_ext or clustered or pair_force" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -15 gpurun_out/r2f_pytest.log
for mode in cell direct; do
CUBEP3M_B200_PPEXT_DENSE=$mode timeout 1500 python bench.py --workload c1x --steps 5 --warmup 3 --no-cpu --evolve-to-z 2.0 --evolve-max-steps 3000 > gpurun_out/r2f_bench_c1x_z2_$mode.json 2> gpurun_out/r2f_bench_c1x_z2_$mode.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2f_bench_c1x_z2_$mode.json'))
print('$mode', d['ms_per_step'], d['config']['evolved'], d['config']['ppext_blocks_tiled_fallback'], d['stage_ms_last_step'])
for k,v in d['stages'].items():
    if k.startswith('pp'): print(k, round(v['ms_per_step'],3), v.get('frac_of_fp32_peak'), v.get('pairs_per_s'), v.get('ordered_pairs_per_step'))
PY
done
