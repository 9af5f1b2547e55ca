"""Print the handful of metrics we read from an `ncu --set full` report exported with
`ncu -i X.ncu-rep --page raw --csv > X.csv`.  usage: python profiles/ncu_metrics.py X.csv"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_op_read.sum',
        'lts__t_sectors_op_write.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('-' * 100)
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:82s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
