#!/bin/bash
# round-2 final evidence, one GPU: the full -m gpu suite, the default bench line (configs[2], with the CPU baseline leg), the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -14 gpurun_out/r2f_pytest.log
/usr/bin/time -v -o gpurun_out/r2f_bench_time.txt timeout 900 python bench.py > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r2f_bench_time.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_c2.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'] and d['cpu_baseline']['ms_per_step'], d['cic_power'] and d['cic_power'].get('ms'), d.get('halofind_peaks'))
PY
/usr/bin/time -v -o gpurun_out/r2f_ref_time.txt timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_c2_reference.json 2> gpurun_out/r2f_bench_c2_reference.err; echo "reference rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r2f_ref_time.txt
cut -c1-600 gpurun_out/r2f_bench_c2_reference.json
free -g | head -2; nproc
