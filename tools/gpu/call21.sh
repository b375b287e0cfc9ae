#!/bin/bash
# wave-aware chunk counts: chunk tests, C2 + C4 (N = 1) and C3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -rf --tb=line -p no:cacheprovider -k "chunk or bench_shape or full_size or padding or golden or long_chains" > $O/c21_pytest.log 2>&1; tail -n 4 $O/c21_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/c21_bench_c2c4.json 2> $O/c21_bench_c2c4.err
timeout 600 python bench.py --steps 5 --warmup 3 --config C3 --no-cpu-baseline --no-e2e --no-c4 > $O/c21_bench_c3.json 2> $O/c21_bench_c3.err
python - <<PY
import json
b=json.loads(open("$O/c21_bench_c2c4.json").read().strip().splitlines()[-1])
print("C2 ms/step", round(b["ms_per_step"],3), "C4", (b.get("strong_c4") or {}).get("ms_per_step"), (b.get("strong_c4") or {}).get("error"))
b=json.loads(open("$O/c21_bench_c3.json").read().strip().splitlines()[-1])
print("C3 ms/step", round(b["ms_per_step"],3), {k:v["ms_per_sweep"] for k,v in list(b["kernels"].items())[:8]}, b["chunk_diagnostics"])
PY
