#!/bin/bash
# 2-GPU checks: sharded sweep parity, sharded fit_model / apply_model, and the two-GPU pytest cases (logs kept).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/c2_smi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 tools/dist_check.py > $O/c2_dist_check.log 2>&1; echo "rc=$?" >> $O/c2_dist_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/dist_fit_check.py > $O/c2_dist_fit_check.log 2>&1; echo "rc=$?" >> $O/c2_dist_fit_check.log
timeout 600 python -m pytest tests -m gpu -q -rs -k "two_gpu or discrete_stateseqs_time_chunks" > $O/c2_pytest_two_gpu.log 2>&1; echo "rc=$?" >> $O/c2_pytest_two_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c2_bench_n2.json 2> $O/c2_bench_n2.err
tail -3 $O/c2_dist_check.log $O/c2_dist_fit_check.log $O/c2_pytest_two_gpu.log
