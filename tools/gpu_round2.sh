#!/bin/bash
# GPU visit: parity suite, c1 + c2 bench, ncu of the x c2r pass and the tiled PP_EXT kernel
mkdir -p gpurun_out
T=${1:-s5c}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -n 4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_c1.json 2> gpurun_out/${T}_c1.err
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err
python - <<PY
import json
for m in ("c1","c2"):
    try:
        d=json.load(open(f"gpurun_out/${T}_{m}.json"))
        st=d["stages"]
        print(m, round(d["ms_per_step"],3), {k: round(st[k]["us_per_launch"],1) for k in ("fft_x_c2r","fft_inv_z_mul","fft_inv_y","fft_x_r2c","fft_fwd_strided","ngp_kick","ppext") if k in st})
    except Exception as e:
        print(m, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_z_sandwich2|fft_strided2" -s 8 -c 3 -o gpurun_out/${T}_fine python bench.py --steps 1 --no-cpu --no-profile > gpurun_out/${T}_ncu1.log 2>&1
ncu -i gpurun_out/${T}_fine.ncu-rep --page raw --csv > gpurun_out/${T}_fine_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppext_tiled -c 1 -o gpurun_out/${T}_ppext python bench.py --workload c0x --steps 1 --no-cpu --no-profile > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_ppext.ncu-rep --page raw --csv > gpurun_out/${T}_ppext_raw.csv 2>/dev/null
tail -n 2 gpurun_out/${T}_ncu1.log gpurun_out/${T}_ncu2.log
