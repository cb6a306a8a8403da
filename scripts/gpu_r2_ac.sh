#!/bin/bash
# round 2, call AC: node-blocked SpMM / Chebyshev step for the modal path
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_modal.py -m gpu -x -q > gpurun_out/pytest_ac.log 2>&1; echo "pytest modal rc=$?"; tail -3 gpurun_out/pytest_ac.log
timeout 600 python bench.py --full-solve 0 --extras 0 --no-cpu-baseline > gpurun_out/bench_ac.json 2> gpurun_out/bench_ac.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_ac.json").read().strip().splitlines()[-1])
print("modal", {k:v for k,v in d["modal"].items() if k!="eigenvalues"})
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_spmm_b2' -s 50 -c 1 \
  -o gpurun_out/prof_r02ac_cheb -f python bench.py --full-solve 0 --extras 0 --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_ac.log 2>&1; echo "ncu rc=$?"
