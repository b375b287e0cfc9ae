#!/bin/bash
# overlap of the scales / observation records with the discrete-state kernels; backprep occupancy variant; warm-up sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "continuous or full_sweep or bench_shape or golden or graph or padding or philox" > $O/c12_pytest.log 2>&1; tail -3 $O/c12_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-c4 > $O/c12_bench_$name.json 2> $O/c12_bench_$name.err
  python - <<PY
import json
try:
    b=json.loads(open("$O/c12_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(b["ms_per_step"],3), {k:v["ms_per_sweep"] for k,v in list(b["kernels"].items())[:3]})
except Exception as e:
    print("$name", "ERR", e)
PY
}
run overlap KPMS_OVERLAP=1
run nooverlap KPMS_OVERLAP=0
run bp4x2 KPMS_BP_CFG=4x2
timeout 200 python tools/warmup_experiment.py > $O/c12_warmup.txt 2>&1; cat $O/c12_warmup.txt | cut -c1-400
