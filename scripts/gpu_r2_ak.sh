#!/bin/bash
# round 2, call AK (gpurun --gpus 2): 2-rank bench after the launch-overhead changes
set -u
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --modal 0 --full-solve 0 > gpurun_out/bench_ak_g2.json 2> gpurun_out/bench_ak_g2.err; echo "bench g2 rc=$?"
python scripts/show_bench.py gpurun_out/bench_ak_g2.json
timeout 300 python bench.py --full-solve 0 --modal 0 --extras 0 --no-cpu-baseline > gpurun_out/bench_ak_g1.json 2> gpurun_out/bench_ak_g1.err; python scripts/show_bench.py gpurun_out/bench_ak_g1.json
