#!/bin/bash
# compute-sanitizer memcheck on the kernels added in round 2 (two-warp filter, two-stage backprep with bulk copies,
# staged observation records, wide-state HMM kernels), small cohort, eager launches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
KPMS_GRAPH=0 timeout 75 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_sweep.py --recordings 2 --frames 500 --sweeps 1 --d 12 > $O/c28_memcheck_d12.log 2>&1; tail -n 3 $O/c28_memcheck_d12.log | cut -c1-200
KPMS_GRAPH=0 timeout 75 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_sweep.py --recordings 2 --frames 500 --sweeps 1 --K 200 > $O/c28_memcheck_K200.log 2>&1; tail -n 3 $O/c28_memcheck_K200.log | cut -c1-200
