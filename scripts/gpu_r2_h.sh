#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tet.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_tet.log 2>&1; echo "pytest tet rc=$?"; tail -5 gpurun_out/pytest_tet.log
timeout 200 python scripts/bench_tet.py > gpurun_out/bench_tet.json 2> gpurun_out/bench_tet.err; cat gpurun_out/bench_tet.json; tail -3 gpurun_out/bench_tet.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tet_assemble' -s 2 -c 1 \
  -o gpurun_out/prof_tet2 -f python scripts/bench_tet.py > gpurun_out/ncu_tet.log 2>&1; echo "ncu tet rc=$?"
