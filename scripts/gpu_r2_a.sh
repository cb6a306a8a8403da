#!/bin/bash
# round 2, call A: the whole GPU test suite on the hardened build + ncu captures that the next kernels need
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log
timeout 200 python scripts/bench_tet.py > gpurun_out/bench_tet.json 2> gpurun_out/bench_tet.err; cat gpurun_out/bench_tet.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tet_assemble' -s 2 -c 1 \
  -o gpurun_out/prof_tet -f python scripts/bench_tet.py > gpurun_out/ncu_tet.log 2>&1; echo "ncu tet rc=$?"
