#!/bin/bash
mkdir -p gpurun_out
for m in main overlap; do
CUBEP3M_B200_MARGIN=$m timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2k_bench_c2_$m.json 2> gpurun_out/r2k_bench_c2_$m.err; echo "bench $m rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench_c2_$m.json'))
print('$m', d['ms_per_step'], d['stage_ms_last_step'])
PY
done
