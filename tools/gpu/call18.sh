#!/bin/bash
# 2 GPUs: bench.py end-to-end number at N = 2 with the current build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > $O/c18_bench_n2.json 2> $O/c18_bench_n2.err
python - <<PY
import json
b=json.loads(open("$O/c18_bench_n2.json").read().strip().splitlines()[-1])
print("ms/step", round(b["ms_per_step"],3), "e2e", b["e2e"], "value", b["value"])
PY
tail -n 3 $O/c18_bench_n2.err | cut -c1-300
