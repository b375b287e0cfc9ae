#!/bin/bash
# backward preparation v2 (prefetched inputs, packed FMAs): parity subset, bench, ncu of the kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "continuous or full_sweep or bench_shape or golden or graph or padding" > $O/c11_pytest.log 2>&1; tail -3 $O/c11_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-c4 > $O/c11_bench.json 2> $O/c11_bench.err
python - <<PY
import json
b=json.loads(open("$O/c11_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(b["ms_per_step"],3), {k:v["ms_per_sweep"] for k,v in list(b["kernels"].items())[:6]})
PY
KPMS_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kalman_backprep_split" -s 1 -c 1 -f -o $O/r02_prof_backprep_v3 python tools/run_sweep.py --recordings 40 --frames 10000 --sweeps 2 > $O/c11_ncu.log 2>&1
tail -2 $O/c11_ncu.log
