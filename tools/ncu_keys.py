"""Key metrics of the first kernel in an .ncu-rep (ncu -i ... --page raw --csv on stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
vals = rows[2] if len(rows) > 2 else rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_cbu.sum',
        'sm__inst_executed_pipe_adu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in hdr:
        print('%-85s %s' % (w, vals[hdr.index(w)]))
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') or 'pipe' in h and 'pct_of_peak_sustained_active' in h and 'inst_executed' in h:
        try:
            if float(vals[i]) >= 0.15:
                print('%-85s %s' % (h.replace('smsp__average_warps_issue_stalled_', 'stall '), vals[i]))
        except ValueError:
            pass
