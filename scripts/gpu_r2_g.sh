#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --full-solve 0 --no-cpu-baseline --pcg-iters 4 --modal 0 > gpurun_out/ncu_launch.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/launches.csv") if l.startswith(chr(34)))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:60]].append(float(r[vi].replace(",","")))
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print(f"{k:62s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us total={sum(v)/1e6:8.3f} ms")
PY
