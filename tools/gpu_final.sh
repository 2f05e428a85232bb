#!/bin/bash
# end-of-session evidence run (1 GPU): full bench lines (c1 with the CPU baseline leg, c2), the ncu launch list of the bench command,
# and one --set full capture of every fine-mesh kernel of a steady-state tile plus the PP_EXT kernels
T=${1:-r1s5}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
timeout 900 python bench.py --workload c2 --steps 5 --no-cpu > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err
CUBEP3M_B200_TILE_STREAMS=2 timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_c1_2streams.json 2> gpurun_out/${T}_bench_c1_2streams.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_c1_reference.json 2> gpurun_out/${T}_bench_c1_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_c1.csv python bench.py --steps 1 --no-cpu --no-profile > gpurun_out/${T}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fft_x_r2c_ngp2|fft_strided2|fft_z_sandwich2|fft_x_c2r3_v4|ngp_kick_kernel" -s 40 -c 12 -o gpurun_out/${T}_fine python bench.py --steps 1 --no-cpu --no-profile > gpurun_out/${T}_ncu_fine.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ppext|ppint" -c 3 -o gpurun_out/${T}_pp python bench.py --workload c0x --steps 1 --no-cpu --no-profile > gpurun_out/${T}_ncu_pp.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/${T}_sanitizer_memcheck.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/${T}_sanitizer_racecheck.txt 2>&1
tail -n 2 gpurun_out/${T}_sanitizer_memcheck.txt gpurun_out/${T}_sanitizer_racecheck.txt
python - <<PY
import json
for m in ("c1","c2","c1_2streams","c1_reference"):
    try:
        d=json.load(open(f"gpurun_out/${T}_bench_{m}.json"))
        print(m, round(d["ms_per_step"],3), d.get("e2e",{}).get("ms_per_step"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("ms_per_step"))
    except Exception as e:
        print(m, "failed", e)
PY
tail -n 2 gpurun_out/${T}_ncu_fine.log gpurun_out/${T}_ncu_pp.log
