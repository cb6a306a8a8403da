#!/bin/bash
# round 2: multi-GPU check of the persistent PCG kernel (run under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout ${T:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
for args in "96 64" "301 77" "1024 512"; do
  run tests/dist_gpu_worker.py $args > gpurun_out/dist_worker_${N}.log 2>&1; echo "dist worker ($args) rc=$?"; grep -E "DIST-OK|Error|error|assert" gpurun_out/dist_worker_${N}.log | head -3
done
run tests/dist_gpu_worker.py 301 177 mag > gpurun_out/dist_worker_${N}_mag.log 2>&1; echo "dist worker (mag) rc=$?"; grep -E "DIST-OK" gpurun_out/dist_worker_${N}_mag.log | tail -1
for mode in persist legacy; do
  if [[ $mode == legacy ]]; then export FE_B200_NO_PERSIST=1; else unset FE_B200_NO_PERSIST; fi
  FE_B200_PERSIST_PROF=1 run bench.py --gpus $N --steps 3 --warmup 3 --full-solve ${FULL:-0} --no-cpu-baseline --modal 0 > gpurun_out/bench_g${N}_$mode.json 2> gpurun_out/bench_g${N}_$mode.err; echo "[$mode] bench s16m rc=$?"
  grep -E "rank 0 grid" gpurun_out/bench_g${N}_$mode.err | tail -2; grep -E "Error|error" gpurun_out/bench_g${N}_$mode.err | head -3
  python scripts/show_bench.py gpurun_out/bench_g${N}_$mode.json
done
