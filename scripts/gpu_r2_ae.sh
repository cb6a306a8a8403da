#!/bin/bash
# round 2, call AE (gpurun --gpus 8): bench lines at 8 ranks with the final build
set -u
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
FE_B200_PERSIST_PROF=1 run bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --modal 0 > gpurun_out/bench_ae_g8.json 2> gpurun_out/bench_ae_g8.err; echo "bench g8 rc=$?"
python scripts/show_bench.py gpurun_out/bench_ae_g8.json
grep -E "rank 0 grid" gpurun_out/bench_ae_g8.err | tail -1 | cut -c1-200
FE_B200_NO_PERSIST=1 run bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --modal 0 --full-solve 0 > gpurun_out/bench_ae_g8_3k.json 2> gpurun_out/bench_ae_g8_3k.err; echo "bench g8 three-kernel rc=$?"
python scripts/show_bench.py gpurun_out/bench_ae_g8_3k.json
run bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --modal 0 --full-solve 0 --nx 1024 --ny 512 > gpurun_out/bench_ae_g8_s1m.json 2> gpurun_out/bench_ae_g8_s1m.err; echo "bench g8 s1m rc=$?"
python scripts/show_bench.py gpurun_out/bench_ae_g8_s1m.json
