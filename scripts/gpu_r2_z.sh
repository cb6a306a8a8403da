#!/bin/bash
# round 2, call Z (gpurun --gpus N): distributed parity worker + bench line at N ranks with the final kernels
set -u
N=${1:-4}
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
for args in "301 77" "301 177 mag" "24 12 tet"; do
  run tests/dist_gpu_worker.py $args > gpurun_out/dist_worker_z_${N}.log 2>&1; echo "dist worker ($args) rc=$?"; grep -E "DIST-OK|Error|error|assert" gpurun_out/dist_worker_z_${N}.log | head -3
done
run bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --modal 0 > gpurun_out/bench_z_g$N.json 2> gpurun_out/bench_z_g$N.err; echo "bench g$N rc=$?"
python scripts/show_bench.py gpurun_out/bench_z_g$N.json
python - $N <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench_z_g{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("config", {k:v for k,v in d["config"].items() if k!="workload"})
print("solve", d.get("solve"))
PY
