#!/bin/bash
# round-2 evidence: launch list of the bench command + ncu --set full summaries of the hot kernels (reports stay on the box, text summaries come back)
mkdir -p gpurun_out
R=/tmp/ncu_reps; mkdir -p $R
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_final_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-profile > gpurun_out/r2f_ncu1.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_summary.py launches gpurun_out/r2_final_launches_c2.csv "bench.py --steps 2 --warmup 3 (default workload = configs[2]); all launches of the run" > gpurun_out/r2_final_launch_summary_c2.txt 2>&1
K1='regex:(fft_z_sandwich2|fft_x_c2r3_v4|fft_x_r2c_ngp2|fft_strided2|ngp_kick_kernel|key_hist_kernel|scatter_kernel|scan_apply_kernel|cic_mass_smem_kernel|pass_pack_kernel)'
timeout 900 ncu --set full --clock-control none -k "$K1" -s 1180 -c 20 -o $R/r2_full_fine python bench.py --steps 1 --warmup 3 --no-cpu --no-profile > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu fine rc=$?"
python tools/ncu_summary.py full $R/r2_full_fine.ncu-rep "ncu --set full, bench.py default workload (configs[2]), first launches of the 4th step" > gpurun_out/r2_final_ncu_full_fine_and_particle_kernels_c2.txt 2>&1
K2='regex:(ppext_tiled_kernel|ppext_margin_roles_kernel|ppext_margin_list_kernel|cic_kick_compact_kernel)'
timeout 900 ncu --set full --clock-control none -k "$K2" -s 12 -c 4 -o $R/r2_full_pp python bench.py --steps 1 --warmup 3 --no-cpu --no-profile > gpurun_out/r2f_ncu3.log 2>&1; echo "ncu pp rc=$?"
python tools/ncu_summary.py full $R/r2_full_pp.ncu-rep "ncu --set full, bench.py default workload (configs[2]): PP_EXT tiled kernel, margin limiter, coarse kick + compaction" > gpurun_out/r2_final_ncu_full_pp_kernels_c2.txt 2>&1
ls -la gpurun_out/ | tail -20; du -sh gpurun_out
ls -la gpurun_out/ | grep r2_final
