import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],stdout=subprocess.PIPE,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]; u=rows[1]; v=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__thread_inst_executed_per_inst_executed.ratio','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.sum.pct','sm__inst_executed_pipe_alu.sum.pct','sm__inst_executed_pipe_lsu.sum.pct','launch__shared_mem_per_block_dynamic','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64']
for i,n in enumerate(h):
    if any(n==w or n.startswith(w) for w in want) and 'per_second' not in n: print(n, u[i], v[i])
print('--- stall reasons (warps per issue)')
for i,n in enumerate(h):
    if 'smsp__average_warps_issue_stalled' in n and n.endswith('.ratio'): print(' ', n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v[i])
