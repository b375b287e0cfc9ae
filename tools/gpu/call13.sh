#!/bin/bash
# wider (latent_dim, nlags) grid + wide-state kernels + async checkpoints: full GPU suite; backprep lockstep variants;
# new kernel-sweep rows; checkpoint overlap
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -rf --tb=line -p no:cacheprovider > $O/c13_pytest.log 2>&1 ) 2> $O/c13_pytest_time.txt
tail -n 40 $O/c13_pytest.log | cut -c1-300
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-c4 > $O/c13_bench_$name.json 2> $O/c13_bench_$name.err
  python - <<PY
import json
try:
    b=json.loads(open("$O/c13_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(b["ms_per_step"],3), {k:v["ms_per_sweep"] for k,v in list(b["kernels"].items())[:3]})
except Exception as e:
    print("$name", "ERR", e)
PY
}
run default KPMS_X=0
run bp4x3s KPMS_BP_CFG=4x3s
run bp6x2s KPMS_BP_CFG=6x2s
run bp12x1s KPMS_BP_CFG=12x1s
run bp12x1 KPMS_BP_CFG=12x1
timeout 400 python tools/kernel_sweep.py --new --reps 3 > $O/c13_kernel_sweep.jsonl 2> $O/c13_kernel_sweep.err; cut -c1-260 $O/c13_kernel_sweep.jsonl; tail -n 3 $O/c13_kernel_sweep.err
timeout 300 python tools/checkpoint_overlap.py --iters 20 --every 5 > $O/c13_checkpoint.jsonl 2> $O/c13_checkpoint.err; cat $O/c13_checkpoint.jsonl; tail -n 3 $O/c13_checkpoint.err
