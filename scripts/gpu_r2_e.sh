#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python scripts/debug_persist.py | grep -v "e-1[3-7]$\|0.000e+00" | head; echo DEBUG-DONE
export FE_B200_PERSIST_PROF=1
timeout 300 python bench.py --steps 3 --warmup 3 --full-solve 0 --no-cpu-baseline --modal 0 > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err; grep "rank 0" gpurun_out/bench_prof.err | tail -1
python scripts/show_bench.py gpurun_out/bench_prof.json | head -1
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run bench.py --gpus 2 --steps 3 --warmup 3 --full-solve 1 --no-cpu-baseline --modal 0 > gpurun_out/bench_g2_persist.json 2> gpurun_out/bench_g2_persist.err; echo "bench g2 rc=$?"
grep -E "rank 0 grid" gpurun_out/bench_g2_persist.err | tail -2
python scripts/show_bench.py gpurun_out/bench_g2_persist.json
