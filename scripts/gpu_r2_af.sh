#!/bin/bash
# round 2, call AF: streamed-SpMV test, modal section timers
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streamed_spmv or s1m or baseline_sizes" > gpurun_out/pytest_af.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_af.log
FE_B200_MODAL_PROF=1 timeout 300 python - <<'PY'
import sys, time, json
sys.path.insert(0, ".")
import bench, torch
for rep in range(2):
    t0 = time.perf_counter()
    r = bench.run_modal(10, 1024, 512, 0)
    print("modal seconds", r["seconds"], "iterations", r["iterations"], "wall", time.perf_counter() - t0, "prof", r.get("prof"))
PY
