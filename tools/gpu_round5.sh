#!/bin/bash
# GPU visit: parity suite, c1 bench, ncu (full) of the x c2r pass and the scan kernels
mkdir -p gpurun_out
T=${1:-s5i}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -n 4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_c1.json 2> gpurun_out/${T}_c1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_c1.json"))
st=d["stages"]
print("c1", round(d["ms_per_step"],3), {k: round(st[k]["us_per_launch"],1) for k in ("fft_x_c2r","fft_inv_z_mul","fft_inv_y","fft_x_r2c","fft_fwd_strided","ngp_kick","scan","key_hist","scatter") if k in st}, d["stage_ms_last_step"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_x_c2r3_v4|scan_reduce|scan_apply" -s 3 -c 4 -o gpurun_out/${T}_k python bench.py --steps 1 --no-cpu --no-profile > gpurun_out/${T}_ncu1.log 2>&1
ncu -i gpurun_out/${T}_k.ncu-rep --page raw --csv > gpurun_out/${T}_k_raw.csv 2>/dev/null
tail -n 2 gpurun_out/${T}_ncu1.log
