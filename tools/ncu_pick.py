"""Print selected metrics of `ncu -i report --page raw --csv` (one kernel per row).
usage: python tools/ncu_pick.py raw.csv [substring ...]"""
import csv
import sys

DEFAULT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
           'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct',
           'sm__warps_active.avg.pct_of_peak', 'launch__registers', 'launch__occupancy_limit', 'launch__grid_size',
           'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_', 'sm__pipe_fp64',
           'l1tex__data_pipe_lsu_wavefronts', 'l1tex__t_requests', 'l1tex__t_sectors', 'l1tex__lsu_writeback',
           'l1tex__data_bank_conflicts', 'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency',
           'l1tex__average_t_sectors_per_request', 'l1tex__f_wavefronts', 'l1tex__m_', 'lts__t_sector_hit_rate',
           'l1tex__t_sector_hit_rate', 'smsp__warps_issue_stalled']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2:] or DEFAULT
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("==", vals[hdr.index("Kernel Name")][:90])
        for h, u, v in zip(hdr, units, vals):
            if any(w in h for w in want):
                print("  %-100s %-12s %s" % (h.split('.', 2)[-1] if h.count('.') > 2 else h, u, v))


if __name__ == "__main__":
    main()
