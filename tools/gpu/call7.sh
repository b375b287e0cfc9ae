#!/bin/bash
# new bench.py: default N=1 line (C2 headline + C4 strong section + CPU port), reference arm, C3, cold start
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/c7_bench_default.json 2> $O/c7_bench_default.err ) 2> $O/c7_time_default.txt
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/c7_bench_reference.json 2> $O/c7_bench_reference.err ) 2> $O/c7_time_reference.txt
timeout 600 python bench.py --steps 10 --warmup 3 --config C3 --no-cpu-baseline --no-e2e > $O/c7_bench_C3.json 2> $O/c7_bench_C3.err
timeout 600 python bench.py --steps 20 --start cold --no-cpu-baseline > $O/c7_bench_cold.json 2> $O/c7_bench_cold.err
timeout 600 python bench.py --steps 20 --warmup 5 --variant states_only --no-cpu-baseline --no-e2e > $O/c7_bench_states_only.json 2> $O/c7_bench_so.err
tail -2 $O/c7_time_default.txt $O/c7_time_reference.txt; tail -3 $O/c7_bench_default.err
