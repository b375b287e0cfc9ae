#!/bin/bash
# 8-GPU bench: weak-scaling headline (C2 per GPU) + strong-scaling C4 section
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 20 --warmup 5 > $O/c8_bench_n8.json 2> $O/c8_bench_n8.err
tail -c 1500 $O/c8_bench_n8.json
