#!/bin/bash
# GPU visit: parity suite with the new defaults (two tiles in flight, single-pass scan), c1 / c2 bench, scan A/B
mkdir -p gpurun_out
T=${1:-s5j}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -n 4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_c1.json 2> gpurun_out/${T}_c1.err
CUBEP3M_B200_SCAN=3pass timeout 600 python bench.py --no-cpu > gpurun_out/${T}_c1_3pass.json 2> gpurun_out/${T}_c1_3pass.err
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err
python - <<PY
import json
for m in ("c1","c1_3pass","c2"):
    try:
        d=json.load(open(f"gpurun_out/${T}_{m}.json"))
        st=d["stages"]
        print(m, round(d["ms_per_step"],3), {k: (round(st[k]["ms_per_step"],3), st[k]["launches_per_step"]) for k in ("scan","key_hist","scatter","ppext") if k in st}, d["stage_ms_last_step"])
    except Exception as e:
        print(m, "failed", e)
PY
