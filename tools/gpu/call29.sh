#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) on the two-warp filter and the lockstep backward preparation
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
KPMS_GRAPH=0 timeout 100 compute-sanitizer --tool racecheck --print-limit 5 python tools/run_sweep.py --recordings 2 --frames 400 --sweeps 1 --d 12 > $O/c29_racecheck_d12.log 2>&1; tail -n 4 $O/c29_racecheck_d12.log | cut -c1-300
