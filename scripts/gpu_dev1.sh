#!/bin/bash
# Development call on a 1-GPU box: GPU tests + quick bench (+ modal timing).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline --modal ${MODAL:-10} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python scripts/show_bench.py gpurun_out/bench_quick.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_quick.json"))
    print("modal", json.dumps(d.get("modal")))
except Exception as e:
    print("no bench json", e)
PY
