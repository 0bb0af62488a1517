#!/bin/bash
# Final single-GPU evidence run (gpurun): tests, both bench arms, the ncu launch list of the bench command and one `ncu --set full`
# capture per hot kernel.  The .ncu-rep files exceed gpurun's 64 MiB pull limit, so each one is turned into its raw-page CSV
# (what tools/ncu_lsu.py / ncu_summary.py / ncu_traffic.py read) on the box and deleted.
set -u
O=gpurun_out
STAGE=${1:-all}
if [ "$STAGE" = all ] || [ "$STAGE" = bench ]; then
python -m pytest tests -m gpu -x -q > $O/fin_tests.log 2>&1; echo rc=$? >> $O/fin_tests.log
python bench.py > $O/fin_bench.json 2> $O/fin_bench.err; echo rc=$? >> $O/fin_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/fin_ref.json 2> $O/fin_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu > $O/fin_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/fin_ncu_list.log 2>&1
fi
if [ "$STAGE" = all ] || [ "$STAGE" = ncu ]; then
python tools/prof_step.py 2 > $O/fin_prof_plain.log 2>&1 || exit 0
cap() {
  ncu --set full --clock-control none -k regex:"$1" --launch-skip $2 -c $3 -o /tmp/fin_$4 python tools/prof_step.py 1 > $O/fin_ncu_$4.log 2>&1
  ncu -i /tmp/fin_$4.ncu-rep --page raw --csv > $O/fin_$4.csv 2>/dev/null
  ls -la /tmp/fin_$4.ncu-rep >> $O/fin_sizes.txt; rm -f /tmp/fin_$4.ncu-rep
}
cap 'k_project_t' 0 1 project
cap 'k_bottom3' 0 2 bottom
cap 'k_knn_f32' 0 1 knn
cap 'k_top_hist' 0 1 hist
cap 'k_top_compact_lean' 5 1 compact
cap 'k_top_relabel_hist' 5 1 relabel
cap 'k_top_finish_warp' 5 1 finish
cap 'k_top_scatter_lean' 0 1 scatter
fi
du -sh $O >> $O/fin_sizes.txt
