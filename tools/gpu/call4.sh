#!/bin/bash
# two-stage backward preparation: parity, then A/B timing against the one-stage kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 600 > $O/c4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c4_pytest.log
tail -15 $O/c4_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/c4_bench_split.json 2> $O/c4_bench_split.err
KPMS_BACKPREP=rows2 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/c4_bench_rows2.json 2> $O/c4_bench_rows2.err
timeout 300 python bench.py --steps 10 --warmup 3 --config C1 --no-cpu-baseline --no-e2e > $O/c4_bench_C1.json 2> $O/c4_bench_C1.err
timeout 300 python bench.py --steps 10 --warmup 3 --config C1 --variant ar_only --no-cpu-baseline --no-e2e > $O/c4_bench_C1_ar.json 2> $O/c4_bench_C1_ar.err
