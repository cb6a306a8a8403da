#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","gpu_launches"): print(k, d[k])
print("roofline", d["roofline"]); print("pcg", d["pcg"]); print("e2e", d["e2e"]); print("solve", d["solve"])
print("modal", {k:v for k,v in d["modal"].items() if k!="eigenvalues"})
print("magnetic", d.get("magnetic")); print("tetrahedra", d.get("tetrahedra")); print("e2e_solve", d.get("e2e_solve"))
print("cpu", {k:v for k,v in d["cpu_baseline"].items() if k!="literal_reference"})
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; cat gpurun_out/bench_ref.json | cut -c1-400
