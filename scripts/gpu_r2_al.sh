#!/bin/bash
# round 2, call AL (gpurun --gpus 8): final build at 8 ranks -- parity worker (tetrahedra: 3x3-block SpMV with halo) + bench line
set -u
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run tests/dist_gpu_worker.py 48 24 tet > gpurun_out/dist_worker_al.log 2>&1; echo "dist worker (48 24 tet) rc=$?"; grep -E "DIST-OK|Error|error|assert" gpurun_out/dist_worker_al.log | head -3
run bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --modal 0 > gpurun_out/bench_al_g8.json 2> gpurun_out/bench_al_g8.err; echo "bench g8 rc=$?"
python scripts/show_bench.py gpurun_out/bench_al_g8.json
