#!/bin/bash
# scales kernel with hoisted constants, transition draw beside the AR draw: full suite, bench with and without the fork
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rf --tb=line -p no:cacheprovider > $O/c23_pytest.log 2>&1; tail -n 5 $O/c23_pytest.log | cut -c1-300
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-c4 > $O/c23_bench_$name.json 2> $O/c23_bench_$name.err
  python - <<PY
import json
try:
    b=json.loads(open("$O/c23_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(b["ms_per_step"],3), [round(x,2) for x in b["step_ms"]], {k:v["ms_per_sweep"] for k,v in b["kernels"].items() if k in ("resample_scales","ar_params","trans_crp","trans_overrides")})
except Exception as e:
    print("$name", "ERR", e)
PY
}
run fork KPMS_PARAM_FORK=1
run nofork KPMS_PARAM_FORK=0
