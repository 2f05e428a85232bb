#!/bin/bash
# clustered-input measurement: configs[1]-size box with PPINT + PP_EXT evolved on the GPU towards z = 1
mkdir -p gpurun_out
timeout 1500 python bench.py --workload c1x --steps 5 --warmup 3 --no-cpu --evolve-to-z ${1:-2.0} --evolve-max-steps 3000 > gpurun_out/r2e_bench_c1x_clustered.json 2> gpurun_out/r2e_bench_c1x_clustered.err; echo "bench rc=$?"
grep evolve gpurun_out/r2e_bench_c1x_clustered.err | tail -40
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_c1x_clustered.json'))
print(d['ms_per_step'], d['config']['evolved'], d['config']['ppext_blocks_tiled_fallback'], d['stage_ms_last_step'])
for k,v in d['stages'].items(): print(k, round(v['ms_per_step'],3), v.get('frac_of_hbm_peak'), v.get('frac_of_fp32_peak'), v.get('pairs_per_s'))
PY
