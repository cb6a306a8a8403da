#!/bin/bash
# round 2, call W: begin_chunk(next) before the store sequence; pipelined tetrahedra (variant 6)
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
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_w.log 2>&1; echo "sanitizer rc=$?"; tail -2 gpurun_out/sanitize_w.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tet.py -m gpu -x -q > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_w.log
for v in 6 5 4; do
  FE_TET_VARIANT=$v timeout 200 python scripts/bench_tet.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tet variant $v', d['assembly_ms'], d['assembly_roofline_frac'], 'pcg', d.get('pcg_ms_per_iter'))"
done
for cfg in "1 1" "0 1" "1 0" "1 1"; do
  set -- $cfg
  ( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_EARLY_BEGIN=$1 -DFE_FAN_EP_LDG=$2" > /dev/null 2>&1 ); echo "EARLY_BEGIN $1 EP_LDG $2"
  for rep in 1 2; do
    timeout 300 python bench.py $B > gpurun_out/bench_w_$1$2_$rep.json 2> gpurun_out/bench_w.err; show gpurun_out/bench_w_$1$2_$rep.json
  done
  timeout 300 python bench.py $B --variant 4 > gpurun_out/bench_w8_$1$2.json 2> gpurun_out/bench_w8.err; show gpurun_out/bench_w8_$1$2.json
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_w_mag_$1$2.json 2> gpurun_out/bench_w_mag.err; show gpurun_out/bench_w_mag_$1$2.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02w_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_w.log 2>&1; echo "ncu rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tet_assemble_pipe' -s 2 -c 1 \
  -o gpurun_out/prof_r02w_tet -f python scripts/bench_tet.py > gpurun_out/ncu_w_tet.log 2>&1; echo "ncu tet rc=$?"
