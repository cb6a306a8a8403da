#!/bin/bash
# round 2, call AG: scalar fan instance with two nodes per lane (64-node chunks)
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "asm ms", round(d["assembly"]["ms"],4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_ag.log 2>&1; echo "sanitizer rc=$?"; tail -2 gpurun_out/sanitize_ag.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/pytest_ag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ag.log
for rep in 1 2; do
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_ag_mag_$rep.json 2> gpurun_out/bench_ag_mag.err; show gpurun_out/bench_ag_mag_$rep.json
done
timeout 300 python bench.py $B --kind magnetic --variant 4 > gpurun_out/bench_ag_mag8.json 2> gpurun_out/bench_ag_mag8.err; show gpurun_out/bench_ag_mag8.json
timeout 300 python bench.py $B --kind magnetic --nx 1024 --ny 512 > gpurun_out/bench_ag_mag_s1m.json 2> gpurun_out/bench_ag_mag_s1m.err; show gpurun_out/bench_ag_mag_s1m.json
timeout 300 python bench.py $B > gpurun_out/bench_ag.json 2> gpurun_out/bench_ag.err; show gpurun_out/bench_ag.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02ag_mag -f python bench.py $B --steps 1 --warmup 3 --kind magnetic > gpurun_out/ncu_ag.log 2>&1; echo "ncu rc=$?"
