"""LSU data-pipe view of an .ncu-rep (kernels bound by shared-memory gathers): wavefronts by source, pipe utilisation, issue, stalls.
Usage: python tools/ncu_lsu.py file.ncu-rep"""
import csv
import subprocess
import sys

# a .csv argument is the saved output of `ncu -i REP --page raw --csv` (the reports themselves are too big to pull from the GPU box)
out = open(sys.argv[1]).read() if sys.argv[1].endswith('.csv') else subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg', 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for row in r[2:]:
    print('=====', row[h.index('Kernel Name')][:70])
    for w in want:
        if w in h:
            print('   %-75s %s' % (w, row[h.index(w)]))
    st = []
    for i, c in enumerate(h):
        if c.startswith('smsp__average_warps_issue_stalled_') and c.endswith('_per_issue_active.ratio') and 'not_issued' not in c:
            try:
                st.append((float(row[i]), c[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    print('   stalls/issue:', [(round(a, 2), b) for a, b in sorted(st, reverse=True)[:6]])
