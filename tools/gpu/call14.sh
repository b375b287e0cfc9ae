#!/bin/bash
# fixes from call 13 (B entries for latent_dim >= 11, test problems), lockstep backprep default + barrier variants,
# appended npz snapshots, ncu of the new default
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -rf --tb=line -p no:cacheprovider > $O/c14_pytest.log 2>&1 ) 2> $O/c14_pytest_time.txt
tail -n 12 $O/c14_pytest.log | cut -c1-300
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-c4 > $O/c14_bench_$name.json 2> $O/c14_bench_$name.err
  python - <<PY
import json
try:
    b=json.loads(open("$O/c14_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(b["ms_per_step"],3), {k:v["ms_per_sweep"] for k,v in list(b["kernels"].items())[:3]})
except Exception as e:
    print("$name", "ERR", e)
PY
}
run default KPMS_X=0
run bp6x2s2 KPMS_BP_CFG=6x2s2
run bp6x2s3 KPMS_BP_CFG=6x2s3
run bp5x2s KPMS_BP_CFG=5x2s
timeout 300 python tools/checkpoint_overlap.py --iters 40 --every 10 > $O/c14_checkpoint.jsonl 2> $O/c14_checkpoint.err; cat $O/c14_checkpoint.jsonl; tail -n 2 $O/c14_checkpoint.err | cut -c1-200
KPMS_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kalman_backprep_split" -s 1 -c 1 -f -o $O/r02_prof_backprep_v4 python tools/run_sweep.py --recordings 40 --frames 10000 --sweeps 2 > $O/c14_ncu.log 2>&1
tail -n 2 $O/c14_ncu.log
