#!/usr/bin/env bash
# ncu capture of the cell kernel on a small-cell sweep point (order 3, B = 32, bare apply)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c23; mkdir -p $O
timeout -k 5 300 ncu --set full --import-source on --clock-control none -k regex:cell_apply_pipe -s 6 -c 1 -f -o $O/pipe_p3 \
   python tools/sweep.py --points 3:32 --out $O/sweep_ncu.json > $O/ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $O/pipe_p3.ncu-rep 12 > $O/pipe_p3_summary.txt 2>&1
ncu -i $O/pipe_p3.ncu-rep --page source --csv > $O/pipe_p3_source.csv 2>/dev/null
python tools/ncu_roles.py $O/pipe_p3_source.csv 10 > $O/pipe_p3_roles.txt 2>&1
head -24 $O/pipe_p3_summary.txt | cut -c1-150
