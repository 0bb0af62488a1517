"""Print the metrics we track from an .ncu-rep (raw page) -- kernel time, DRAM traffic, occupancy, stalls."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_shfl.sum' if False else 'smsp__inst_executed_pipe_lsu.sum']


def main(path):
    out = open(path).read() if path.endswith('.csv') else subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('=====', r[hdr.index('Kernel Name')][:60], '| id', r[hdr.index('ID')])
        for w in WANT:
            if w in hdr:
                print('   %s: %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
        items = []
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h:
                try:
                    items.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        print('   stalls/issue:', [(round(v, 2), h) for v, h in sorted(items, reverse=True)[:6]])


if __name__ == '__main__':
    main(sys.argv[1])
