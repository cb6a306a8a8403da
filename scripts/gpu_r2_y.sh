#!/bin/bash
# round 2, call Y: TMA L2 prefetch of the coordinates a chunk touches first
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
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_y.log 2>&1; echo "sanitizer rc=$?"; tail -2 gpurun_out/sanitize_y.log
for e in 1 0 1; do
  ( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_L2PF=$e" > /dev/null 2>&1 ); echo "L2PF $e"
  for rep in 1 2; do
    timeout 300 python bench.py $B > gpurun_out/bench_y_${e}_$rep.json 2> gpurun_out/bench_y.err; show gpurun_out/bench_y_${e}_$rep.json
  done
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_y_mag_$e.json 2> gpurun_out/bench_y_mag.err; show gpurun_out/bench_y_mag_$e.json
  timeout 300 python bench.py $B --nx 1024 --ny 512 > gpurun_out/bench_y_s1m_$e.json 2> gpurun_out/bench_y_s1m.err; show gpurun_out/bench_y_s1m_$e.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02y_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_y.log 2>&1; echo "ncu rc=$?"
