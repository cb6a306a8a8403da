#!/bin/bash
# round 2, call AI: lane-level L2 prefetch of the coordinates a chunk touches first (FE_FAN_PF trips ahead)
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
for pf in 2 0 4 1; do
  ( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_PF=$pf" > /dev/null 2>&1 ); echo "FE_FAN_PF $pf"
  timeout 300 python bench.py $B > gpurun_out/bench_ai_$pf.json 2> gpurun_out/bench_ai.err; show gpurun_out/bench_ai_$pf.json
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_ai_mag_$pf.json 2> gpurun_out/bench_ai_mag.err; show gpurun_out/bench_ai_mag_$pf.json
done
