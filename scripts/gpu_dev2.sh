#!/bin/bash
# Development call on a 2-GPU box: GPU tests, 1-GPU quick bench, 2-GPU parity worker + bench (p2p / nccl).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python scripts/show_bench.py gpurun_out/bench_quick.json
MODES="${MODES:-p2p}" FULL=${FULL:-1} bash scripts/gpu_dist_check.sh 2
timeout 600 python bench.py --kind magnetic --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline > gpurun_out/bench_mag.json 2> gpurun_out/bench_mag.err; tail -3 gpurun_out/bench_mag.err
python scripts/show_bench.py gpurun_out/bench_mag.json
timeout 300 python bench.py --kind magnetic --nx 1024 --ny 512 --steps 5 --warmup 3 --full-solve 1 --no-cpu-baseline > gpurun_out/bench_mag_s1m.json 2>> gpurun_out/bench_mag.err
python scripts/show_bench.py gpurun_out/bench_mag_s1m.json
