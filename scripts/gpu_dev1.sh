#!/bin/bash
# Development call on a 1-GPU box: GPU tests + quick plane-stress bench (+ modal timing) + magnetic bench.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline --modal ${MODAL:-10} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python scripts/show_bench.py gpurun_out/bench_quick.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_quick.json"))
    print("modal", json.dumps(d.get("modal"))[:400])
except Exception as e:
    print("no bench json", e)
PY
timeout 600 python bench.py --kind magnetic --steps 5 --warmup 3 --full-solve ${MAGFULL:-1} --no-cpu-baseline > gpurun_out/bench_mag.json 2> gpurun_out/bench_mag.err; tail -3 gpurun_out/bench_mag.err
python scripts/show_bench.py gpurun_out/bench_mag.json
timeout 300 python bench.py --kind magnetic --nx 1024 --ny 512 --steps 5 --warmup 3 --full-solve 1 --no-cpu-baseline > gpurun_out/bench_mag_s1m.json 2>> gpurun_out/bench_mag.err
python scripts/show_bench.py gpurun_out/bench_mag_s1m.json
