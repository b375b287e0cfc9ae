#!/bin/bash
# ncu --set full of the two row-per-lane filters (final build); launches per sweep: main, re-run -> skip 2 = second sweep's main launch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
KPMS_GRAPH=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"kalman_forward_rows_kernel" -s 2 -c 1 -f -o $O/r02_prof_forward_rows python tools/run_sweep.py --recordings 40 --frames 10000 --sweeps 2 > $O/c24_ncu_a.log 2>&1; tail -n 1 $O/c24_ncu_a.log
KPMS_GRAPH=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"kalman_forward_rows2w_kernel" -s 2 -c 1 -f -o $O/r02_prof_forward_rows2w python tools/run_sweep.py --recordings 40 --frames 10000 --sweeps 2 --d 12 > $O/c24_ncu_b.log 2>&1; tail -n 1 $O/c24_ncu_b.log
