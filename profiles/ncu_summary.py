import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum ','dram__bytes_write.sum ','sm__inst_executed.avg.per_cycle_elapsed','smsp__issue_active.avg.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread ','launch__grid_size','launch__block_size','launch__occupancy_limit','smsp__average_warps_issue_stalled','smsp__average_warp_latency_issue_stalled','sm__inst_executed_pipe_alu.avg.pct','sm__inst_executed_pipe_fma.avg.pct','sm__inst_executed_pipe_lsu','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum ','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ','l1tex__t_sector_hit_rate','lts__t_sector_hit_rate','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','smsp__pcsamp_warps_issue_stalled']
for h,u,v in zip(hdr,units,vals):
    hh=h+' '
    if any(w in hh for w in want):
        try:
            if float(v.replace(',',''))==0: continue
        except: pass
        print(f'{h:95s} {u:16s} {v}')
