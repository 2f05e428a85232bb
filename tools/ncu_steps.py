import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','lts__t_sectors_op_write.sum','lts__t_sectors_srcunit_tex_op_write.sum']
for vals in rows[2:]:
    print('---',vals[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f"  {w:68s} {vals[i]} {units[i]}")
    st=[(float(vals[i]),h) for i,h in enumerate(hdr) if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h]
    tot=sum(v for v,_ in st)
    print('  stalls: '+' | '.join(f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio','')} {100*v/tot:.0f}%" for v,h in sorted(st,reverse=True)[:8]))
