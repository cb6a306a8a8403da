#!/bin/bash
# Under `gpurun --gpus N`: distributed parity worker + N-GPU bench lines.
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout ${T:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run tests/dist_gpu_worker.py 96 64 > gpurun_out/dist_worker_$N.log 2>&1; echo "dist worker rc=$?"; tail -3 gpurun_out/dist_worker_$N.log
run tests/dist_gpu_worker.py 301 77 >> gpurun_out/dist_worker_$N.log 2>&1; echo "dist worker(301x77) rc=$?"; tail -1 gpurun_out/dist_worker_$N.log
run bench.py --gpus $N --steps 3 --warmup 3 --nx 1024 --ny 512 --full-solve 1 > gpurun_out/bench_s1m_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench s1m rc=$?"; tail -2 gpurun_out/bench_g$N.err
run bench.py --gpus $N --steps 3 --warmup 3 --full-solve ${FULL:-0} > gpurun_out/bench_g$N.json 2>> gpurun_out/bench_g$N.err; echo "bench s16m rc=$?"; tail -2 gpurun_out/bench_g$N.err
cat gpurun_out/bench_g$N.json
