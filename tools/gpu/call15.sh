#!/bin/bash
# coalesced observation records, squeeze + float32 proposal in the gamma draw: full suite + bench with e2e
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -rf --tb=line -p no:cacheprovider > $O/c15_pytest.log 2>&1 ) 2> $O/c15_pytest_time.txt
tail -n 8 $O/c15_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > $O/c15_bench.json 2> $O/c15_bench.err
python - <<PY
import json
b=json.loads(open("$O/c15_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(b["ms_per_step"],3), "e2e", b["e2e"]["ms_per_step"], {k:v["ms_per_sweep"] for k,v in list(b["kernels"].items())[:12]})
PY
