#!/bin/bash
# round-2 final evidence, one GPU: the default bench line (configs[2], with the CPU baseline leg) and the reference arm
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_c2.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'] and d['cpu_baseline']['ms_per_step'], d['cic_power'] and d['cic_power'].get('ms'), d.get('halofind_peaks'))
PY
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_c2_reference.json 2> gpurun_out/r2f_bench_c2_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - t0 ))s"
cut -c1-700 gpurun_out/r2f_bench_c2_reference.json
