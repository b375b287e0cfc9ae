#!/bin/bash
# 8 GPUs, final build: weak-scaling headline (C2 per GPU) + strong-scaling C4 section
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 10 --warmup 3 > $O/c22_bench_n8.json 2> $O/c22_bench_n8.err ) 2> $O/c22_time.txt
python - <<PY
import json
b=json.loads(open("$O/c22_bench_n8.json").read().strip().splitlines()[-1])
print("N8 ms/step", round(b["ms_per_step"],3), "value", round(b["value"]/1e6,1), "e2e", b["e2e"], "c4", (b.get("strong_c4") or {}).get("ms_per_step"), (b.get("strong_c4") or {}).get("value"))
PY
grep real $O/c22_time.txt; tail -n 2 $O/c22_bench_n8.err | cut -c1-300
