#!/bin/bash
# two-warp row-per-lane filter for 32 < n <= 64 (float32): parity of every pair, chunked n = 36, kernel-sweep rows
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -rf --tb=line -p no:cacheprovider > $O/c16_pytest.log 2>&1 ) 2> $O/c16_pytest_time.txt
tail -n 12 $O/c16_pytest.log | cut -c1-300
timeout 300 python tools/kernel_sweep.py --dims 12,16 --reps 3 > $O/c16_kernel_sweep.jsonl 2> $O/c16_kernel_sweep.err; cut -c1-330 $O/c16_kernel_sweep.jsonl; tail -n 3 $O/c16_kernel_sweep.err
