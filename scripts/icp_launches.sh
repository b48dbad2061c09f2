#!/bin/bash
# per-kernel durations of the ICP iteration (serialised by ncu, cold caches: compare shares / variants, not absolutes)
#   scripts/icp_launches.sh <out.csv> [lib.so]
out=$1; lib=$2
LS3D_B200_LIB=$lib python scripts/prof_icp.py 4 > /dev/null 2>&1 || exit 1
LS3D_B200_LIB=$lib ncu --metrics gpu__time_duration.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:k_icp -c 60 --csv --log-file $out python scripts/prof_icp.py 4 > /dev/null 2>&1
python scripts/icp_launch_table.py $out
