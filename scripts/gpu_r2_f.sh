#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
FE_B200_PERSIST_PROF=1 run bench.py --gpus $N --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline --modal 0 > gpurun_out/bench_g${N}_persist.json 2> gpurun_out/bench_g${N}_persist.err; echo "bench rc=$?"
grep -E "rank [03] grid" gpurun_out/bench_g${N}_persist.err | tail -3
python scripts/show_bench.py gpurun_out/bench_g${N}_persist.json | head -1
