#!/bin/bash
# last evidence refresh (1 GPU): parity suite, bench lines with the final defaults (two fine tiles in flight), launch list
T=${1:-r1s5}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -n 3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
timeout 900 python bench.py --workload c2 --steps 5 --no-cpu > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_c1_reference.json 2> gpurun_out/${T}_bench_c1_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_c1.csv python bench.py --steps 1 --no-cpu --no-profile > gpurun_out/${T}_launches.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -n 1 gpurun_out/${T}_smoke.log
python - <<PY
import json
for m in ("c1","c2","c1_reference"):
    try:
        d=json.load(open(f"gpurun_out/${T}_bench_{m}.json"))
        print(m, round(d["ms_per_step"],3), d.get("e2e",{}).get("ms_per_step"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("ms_per_step"), d.get("gpu_launches"))
    except Exception as e:
        print(m, "failed", e)
PY
