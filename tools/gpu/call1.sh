#!/bin/bash
# Round-2 first GPU call: full GPU test-suite (incl. benchmark-shape parity), HMM refinement validation, benches, micro.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 600 > $O/c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c1_pytest.log
KPMS_HMM_REFINE=3 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "time_chunks or discrete" > $O/c1_pytest_refine.log 2>&1; echo "rc=$?" >> $O/c1_pytest_refine.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/c1_bench_C2.json 2> $O/c1_bench_C2.err
timeout 300 python bench.py --steps 20 --warmup 5 --variant states_only --no-cpu-baseline > $O/c1_bench_C2_states_only.json 2> $O/c1_bench_C2_so.err
timeout 300 python bench.py --steps 20 --warmup 5 --variant ar_only --config C1 --no-cpu-baseline > $O/c1_bench_C1_ar_only.json 2> $O/c1_bench_C1_ar.err
timeout 300 python bench.py --steps 20 --warmup 5 --config C1 --no-cpu-baseline > $O/c1_bench_C1_full.json 2> $O/c1_bench_C1_full.err
timeout 200 python tools/cold_start.py --sweeps 12 > $O/c1_cold_default.jsonl 2>&1
KPMS_HMM_REFINE=3 timeout 200 python tools/cold_start.py --sweeps 12 > $O/c1_cold_refine3.jsonl 2>&1
timeout 60 tools/micro/mma_sync_peak > $O/c1_mma_sync_peak.txt 2>&1
tail -3 $O/c1_pytest.log
