#!/bin/bash
# warm-up length of the verified chunks (bench, graph mode), Kalman per-kernel times at latent_dim 10 / 12 / 16
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-c4 > $O/c19_bench_$name.json 2> $O/c19_bench_$name.err
  python - <<PY
import json
try:
    b=json.loads(open("$O/c19_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(b["ms_per_step"],3), {k:v["ms_per_sweep"] for k,v in b["kernels"].items() if k in ("kalman_forward","hmm_forward","kalman_affine","hmm_backward","kalman_forward_rerun","hmm_forward_rerun","hmm_forward_refine")}, b["chunk_diagnostics"])
except Exception as e:
    print("$name", "ERR", e)
PY
}
run w64 KPMS_WARMUP=64
run w48 KPMS_WARMUP=48
run w32 KPMS_WARMUP=32
timeout 200 python tools/prof_kalman_dims.py 10 12 16 > $O/c19_kalman_dims.jsonl 2> $O/c19_kalman_dims.err; cut -c1-700 $O/c19_kalman_dims.jsonl; tail -n 2 $O/c19_kalman_dims.err
