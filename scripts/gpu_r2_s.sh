#!/bin/bash
# round 2, call S: experiments on the chunk-begin stall of the register-walk kernel
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
export FE_B200_FAN_DESIGN=rw
for e in 0 1 2 3 0; do
  ( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_RW_EXP=$e" > /dev/null 2>&1 ); echo "EXP $e"
  for rep in 1 2; do
    timeout 300 python bench.py $B > gpurun_out/bench_s_e${e}_$rep.json 2> gpurun_out/bench_s_e${e}.err; show gpurun_out/bench_s_e${e}_$rep.json
  done
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_s_mag_e$e.json 2> gpurun_out/bench_s_mag_e$e.err; show gpurun_out/bench_s_mag_e$e.json
done
( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_RW_EXP=3" > /dev/null 2>&1 )
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02s_e3 -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_s_e3.log 2>&1; echo "ncu e3 rc=$?"
( cd finite_elements_b200/csrc && touch assemble.cu && make > /dev/null 2>&1 )
