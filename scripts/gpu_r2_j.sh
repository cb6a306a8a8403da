#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tet.py tests/test_gpu_parity.py -m gpu -x -q -k "tet or magnetic or spmv or solution" > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_j.log
timeout 200 python scripts/bench_tet.py 2>/dev/null
