#!/bin/bash
# launch-shape A/B of the two-stage backward preparation (C2): warps x CTAs per SM, lockstep phases on/off
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "continuous or bench_shape or full_sweep" > $O/c6_pytest.log 2>&1; tail -3 $O/c6_pytest.log
for cfg in default 4x3 6x2 6x2L 12x1L 5x2L; do
  KPMS_BP_CFG=$cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/c6_bench_$cfg.json 2> $O/c6_bench_$cfg.err
  python - <<PY
import json
b=json.loads(open("$O/c6_bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", "ms/step", round(b["ms_per_step"],3), "backprep", b["kernels"]["kalman_backprep"]["ms_per_sweep"])
PY
done
