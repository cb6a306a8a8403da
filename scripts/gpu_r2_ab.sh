#!/bin/bash
# round 2, call AB (gpurun --gpus 2): distributed parity tests with the 3x3-block SpMV (halo variant)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; echo "pytest dist rc=$?"; tail -4 gpurun_out/pytest_ab.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run tests/dist_gpu_worker.py 40 20 tet > gpurun_out/dist_worker_ab.log 2>&1; echo "dist worker (40 20 tet) rc=$?"; grep -E "DIST-OK|Error|error|assert" gpurun_out/dist_worker_ab.log | head -3
FE_B200_NO_P2P=1 run tests/dist_gpu_worker.py 24 12 tet > gpurun_out/dist_worker_ab2.log 2>&1; echo "dist worker nccl (24 12 tet) rc=$?"; grep -E "DIST-OK|Error|error|assert" gpurun_out/dist_worker_ab2.log | head -3
