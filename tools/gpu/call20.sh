#!/bin/bash
# filter walk stops at the last valid frame: full suite, default bench exactly as the driver runs it, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -rf --tb=line -p no:cacheprovider > $O/c20_pytest.log 2>&1 ) 2> $O/c20_pytest_time.txt
tail -n 6 $O/c20_pytest.log | cut -c1-300
( time timeout 900 python bench.py > $O/c20_bench_default.json 2> $O/c20_bench_default.err ) 2> $O/c20_time_default.txt
python - <<PY
import json
b=json.loads(open("$O/c20_bench_default.json").read().strip().splitlines()[-1])
print("ms/step", round(b["ms_per_step"],3), "value", round(b["value"]/1e6,2), "e2e", b["e2e"]["ms_per_step"], "c4", (b.get("strong_c4") or {}).get("ms_per_step"), "cpu", (b.get("cpu_baseline") or {}).get("value"), {k:v["ms_per_sweep"] for k,v in list(b["kernels"].items())[:8]}, b["roofline"]["compute"], b["clocks"])
PY
grep real $O/c20_time_default.txt
