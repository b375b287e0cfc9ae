#!/bin/bash
# ncu --set full capture of the two Kalman factorisation kernels (second sweep of a 40-chain x 10 000-frame cohort)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
KPMS_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kalman_backprep_split|kalman_forward_rows" -s 3 -c 3 -f -o $O/r02_prof_kalman python tools/run_sweep.py --recordings 40 --frames 10000 --sweeps 2 > $O/c5_ncu.log 2>&1
tail -5 $O/c5_ncu.log
ls -la $O/*.ncu-rep
