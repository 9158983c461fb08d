"""Print the metrics we track from `ncu -i X.ncu-rep --page raw --csv` output (run on the CPU box)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'launch__shared_mem_per_block', 'launch__grid_size', 'launch__block_size', 'launch__cluster',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma', 'sm__pipe_fma', 'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_lsu',
        'sm__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_xu', 'sm__inst_executed_pipe_tensor', 'sm__pipe_tensor',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'smsp__issue_active.avg.pct', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct',
        'smsp__average_warp', 'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'smsp__inst_executed_op_shared', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if any(w in h for w in want):
            print(f"{h:100s} {u:14s} {v}")
