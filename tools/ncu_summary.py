import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','sm__cycles_elapsed.max','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('-----')
    for k in keys:
        if k in d: print(f"{k:85s} {d[k][:90]} {rows[1][hdr.index(k)]}")
    st={k:float(v.replace(',','')) for k,v in d.items() if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k and v not in ('','n/a')}
    tot=sum(st.values())
    print('stalls:', ', '.join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_','')} {100*v/tot:.0f}%" for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:8]))
