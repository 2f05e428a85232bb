#!/bin/bash
# PP_EXT tiled kernel: parity tests, A/B bench on c2 (512^3 particles), ncu capture on the small c0x box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ppext_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pp_ext or pair_force or pm_pp_lcdm" > gpurun_out/ppext_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ppext_pytest.log
tail -5 gpurun_out/ppext_pytest.log
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu > gpurun_out/ppext_c2_tiled.json 2> gpurun_out/ppext_c2_tiled.err
CUBEP3M_B200_PPEXT=direct timeout 600 python bench.py --workload c2 --steps 5 --no-cpu > gpurun_out/ppext_c2_direct.json 2> gpurun_out/ppext_c2_direct.err
python - <<'PY'
import json
for m in ("tiled","direct"):
    try:
        d=json.load(open(f"gpurun_out/ppext_c2_{m}.json"))
        print(m, d["ms_per_step"], d["stages"].get("ppext"), d["stage_ms_last_step"])
    except Exception as e:
        print(m, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppext_tiled -c 1 -o gpurun_out/ppext_tiled_c0x python bench.py --workload c0x --steps 1 --no-cpu --no-profile > gpurun_out/ppext_ncu.log 2>&1
ncu -i gpurun_out/ppext_tiled_c0x.ncu-rep --page raw --csv > gpurun_out/ppext_tiled_c0x_raw.csv 2>/dev/null
tail -3 gpurun_out/ppext_ncu.log
