#!/bin/bash
# round 2, call O: fan kernel v4 (all-async inputs, single-fan fast walk, coalesced copy-out)
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_o.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitize_o.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_o.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "asm ms", round(d["assembly"]["ms"],4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], "pcg ms/it", round(d["pcg"]["ms_per_iter"],4))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
run() {  # $1 = tag
  timeout 300 python bench.py $B > gpurun_out/bench_o_$1.json 2> gpurun_out/bench_o_$1.err; show gpurun_out/bench_o_$1.json
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_o_mag_$1.json 2> gpurun_out/bench_o_mag_$1.err; show gpurun_out/bench_o_mag_$1.json
}
run w2
timeout 300 python bench.py $B --variant 4 > gpurun_out/bench_o_v4.json 2> gpurun_out/bench_o_v4.err; show gpurun_out/bench_o_v4.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02o_s16m_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_o_full.log 2>&1; echo "ncu asm rc=$?"
( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_WARPS=1 -DFE_FAN_MINB=10" > /dev/null 2>&1 ); echo "warps/CTA 1"
run w1
( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_WARPS=4 -DFE_FAN_MINB=2" > /dev/null 2>&1 ); echo "warps/CTA 4"
run w4
( cd finite_elements_b200/csrc && touch assemble.cu && make > /dev/null 2>&1 )
