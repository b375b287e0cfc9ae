#!/bin/bash
# 2 GPUs: why the end-to-end step is slow at N > 1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/e2e_probe.py > $O/c17_probe_n2.jsonl 2> $O/c17_probe_n2.err
cat $O/c17_probe_n2.jsonl; tail -n 3 $O/c17_probe_n2.err | cut -c1-300
timeout 200 python tools/e2e_probe.py > $O/c17_probe_n1.jsonl 2> $O/c17_probe_n1.err
cat $O/c17_probe_n1.jsonl; tail -n 2 $O/c17_probe_n1.err | cut -c1-300
nvidia-smi topo -m > $O/c17_topo.txt 2>&1; head -n 8 $O/c17_topo.txt; nproc; free -g | head -2
