#!/bin/bash
# final build: default bench exactly as the driver runs it, smoke, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 python bench.py > $O/c25_bench_default.json 2> $O/c25_bench_default.err ) 2> $O/c25_time.txt
python - <<PY
import json
b=json.loads(open("$O/c25_bench_default.json").read().strip().splitlines()[-1])
print("ms/step", round(b["ms_per_step"],3), "value", round(b["value"]/1e6,2), "e2e", b["e2e"]["ms_per_step"], "c4", (b.get("strong_c4") or {}).get("ms_per_step"), "cpu", (b.get("cpu_baseline") or {}).get("value"), b["roofline"]["compute"]["frac"], b["gpu_launches"], b["clocks"])
PY
grep real $O/c25_time.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
