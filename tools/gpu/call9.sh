#!/bin/bash
# 2-GPU: teardown after graphed sweeps with the all-reduce inside (bench default path incl. C4 section; sharded fit)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 5 --warmup 3 > $O/c9_bench_n2.json 2> $O/c9_bench_n2.err ) 2> $O/c9_time.txt
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/dist_fit_check.py > $O/c9_dist_fit.log 2>&1 ) 2>> $O/c9_time.txt
grep real $O/c9_time.txt; tail -2 $O/c9_dist_fit.log; tail -c 600 $O/c9_bench_n2.json
