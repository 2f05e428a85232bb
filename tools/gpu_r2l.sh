#!/bin/bash
# round-2 evidence: launch list of the default bench command + ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-profile > gpurun_out/r2l_ncu1.log 2>&1; echo "ncu launches rc=$?"
K1='regex:(fft_z_sandwich2|fft_x_c2r3_v4|fft_x_r2c_ngp2|fft_strided2|ngp_kick_kernel|key_hist_kernel|scatter_kernel|scan_apply_kernel|cic_mass_smem_kernel|pass_pack_kernel)'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K1" -s 1180 -c 24 -o gpurun_out/r2_full_fine python bench.py --steps 1 --warmup 3 --no-cpu --no-profile > gpurun_out/r2l_ncu2.log 2>&1; echo "ncu fine rc=$?"
K2='regex:(ppext_tiled_kernel|ppext_margin_roles_kernel|ppext_margin_list_kernel|cic_kick_compact_kernel)'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K2" -s 12 -c 4 -o gpurun_out/r2_full_pp python bench.py --steps 1 --warmup 3 --no-cpu --no-profile > gpurun_out/r2l_ncu3.log 2>&1; echo "ncu pp rc=$?"
timeout 600 python tools/profile_pp_clustered.py 3 > gpurun_out/r2l_clustered.log 2>&1; tail -3 gpurun_out/r2l_clustered.log
K3='regex:(ppext_cell_kernel|ppint_kernel|ppext_tiled_kernel)'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K3" -s 3 -c 3 -o gpurun_out/r2_full_pp_clustered python tools/profile_pp_clustered.py 2 > gpurun_out/r2l_ncu4.log 2>&1; echo "ncu clustered rc=$?"
ls -la gpurun_out/*.ncu-rep
