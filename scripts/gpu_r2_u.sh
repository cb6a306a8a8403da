#!/bin/bash
# round 2, call U (gpurun --gpus 2): distributed parity tests + 2-GPU bench lines with the final kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_u.log 2>&1; echo "pytest dist rc=$?"; tail -4 gpurun_out/pytest_u.log
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --modal 0 > gpurun_out/bench_u_g2.json 2> gpurun_out/bench_u_g2.err; echo "bench g2 rc=$?"
python scripts/show_bench.py gpurun_out/bench_u_g2.json
FE_B200_PERSIST=1 run bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --modal 0 --full-solve 0 > gpurun_out/bench_u_g2_persist.json 2> gpurun_out/bench_u_g2_persist.err; echo "bench g2 persist rc=$?"
python scripts/show_bench.py gpurun_out/bench_u_g2_persist.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_u_g2.json").read().strip().splitlines()[-1])
print("config", {k:v for k,v in d["config"].items() if k!="workload"})
print("solve", d.get("solve"))
PY
