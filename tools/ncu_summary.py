"""Dev tool: compact per-kernel table from `ncu --page raw --csv` output.  Usage: ncu_summary.py file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
names = {
    'gpu__time_duration.sum': 'us', 'dram__bytes_read.sum': 'rdMB', 'dram__bytes_write.sum': 'wrMB',
    'lts__t_sector_hit_rate.pct': 'L2hit', 'l1tex__t_sector_hit_rate.pct': 'L1hit',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'launch__registers_per_thread': 'regs',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1%',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2%', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
    'smsp__inst_executed.sum': 'winst', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'st_long',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'st_short',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio': 'st_lg',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio': 'st_mio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'st_bar',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'st_math',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'st_wait',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active': 'fp64%',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'fp64cyc%',
    'l1tex__data_pipe_lsu_wavefronts.sum': 'l1wf', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smemwf',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'bankconf',
}
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    out = [r[idx['Kernel Name']][:44]]
    for n, short in names.items():
        if n in idx and r[idx[n]] != '':
            try:
                v = float(r[idx[n]])
                out.append(f"{short}={v:.4g}")
            except ValueError:
                pass
    print('  '.join(out))
