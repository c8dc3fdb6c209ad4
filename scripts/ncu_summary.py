"""Print the key ncu metrics (duration, occupancy limiters, DRAM bytes, throughput, stall reasons) of every kernel in an .ncu-rep: python scripts/ncu_summary.py file.ncu-rep"""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines()))
h=r[0]; idx={n:i for i,n in enumerate(h)}
want=['gpu__time_duration.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','launch__waves_per_multiprocessor']
for row in r[2:]:
    print('----', row[idx['Kernel Name']][:90])
    for w in want:
        if w in idx: print('   %-90s %s %s'%(w,row[idx[w]], r[1][idx[w]]))
