#!/bin/bash
# Source-level (SASS) ncu capture of one kernel of the bench workload: writes the source page as CSV (the .ncu-rep is too large to pull).
# usage: tools/src_profile.sh <kernel regex> <launch-skip> <tag>
set -u
O=gpurun_out
python tools/prof_step.py 1 > $O/src_plain_$3.log 2>&1 || exit 0
ncu --set full --import-source on --clock-control none -k regex:"$1" --launch-skip $2 -c 1 -o /tmp/src_$3 python tools/prof_step.py 1 > $O/src_ncu_$3.log 2>&1
ncu -i /tmp/src_$3.ncu-rep --page source --csv > $O/src_$3.csv 2>/dev/null
ncu -i /tmp/src_$3.ncu-rep --page raw --csv > $O/src_$3_raw.csv 2>/dev/null
ls -la /tmp/src_$3.ncu-rep >> $O/src_sizes.txt
