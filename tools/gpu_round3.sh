#!/bin/bash
# short GPU visit: PP parity tests, c2 bench, ncu of the tiled PP_EXT kernel
mkdir -p gpurun_out
T=${1:-s5e}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pp_ext or pair_force or smoke" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -n 4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_c2.json"))
print("c2", round(d["ms_per_step"],3), d["stages"]["ppext"], d["stage_ms_last_step"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppext_tiled -c 1 -o gpurun_out/${T}_ppext python bench.py --workload c0x --steps 1 --no-cpu --no-profile > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_ppext.ncu-rep --page raw --csv > gpurun_out/${T}_ppext_raw.csv 2>/dev/null
tail -n 2 gpurun_out/${T}_ncu2.log
