#!/bin/bash
# CUDA-graph sweep: tests, then bench with and without graphs.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 600 -x > $O/c3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c3_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/c3_bench_graph.json 2> $O/c3_bench_graph.err
KPMS_GRAPH=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/c3_bench_eager.json 2> $O/c3_bench_eager.err
KPMS_BENCH_NO_CLOCKS=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/c3_bench_graph_noclk.json 2> $O/c3_bench_graph_noclk.err
tail -5 $O/c3_pytest.log
