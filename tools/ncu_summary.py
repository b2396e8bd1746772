"""Summarise an `ncu --page raw --csv` dump: one block per profiled launch with the metrics the roofline needs."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct']


def main(path, extra=()):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('---', r[col['Kernel Name']][:90], 'grid', r[col.get('launch__grid_size', 0)])
        for w in list(WANT) + [h for h in hdr if any(e in h for e in extra)]:
            if w in col and r[col[w]] not in ('', 'n/a'):
                print('   %-78s %s %s' % (w, r[col[w]], units[col[w]]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
