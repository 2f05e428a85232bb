#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "drift_link or full_step or pp_ext_tiled_lcdm or overflow or replay or smoke or scan_variants" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
tail -3 gpurun_out/r2w_pytest.log
for w in c1 c2; do
timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu > gpurun_out/r2w_bench_$w.json 2> gpurun_out/r2w_bench_$w.err; echo "bench $w rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2w_bench_$w.json'))
print("$w", d['ms_per_step'], d['stage_ms_last_step'])
print({k: round(v['ms_per_step'],3) for k,v in d['stages'].items() if k in ('key_hist','scan','scatter','misc','ppext','ppext_margin')})
PY
done
