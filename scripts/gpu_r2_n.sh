#!/bin/bash
# round 2, call N: fan kernel v3 (deeper TMA ring, early first gathers), staged tets v2
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_n.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitize_n.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tet.py -m gpu -x -q > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_n.log
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
  timeout 300 python bench.py $B > gpurun_out/bench_n_$1.json 2> gpurun_out/bench_n_$1.err; show gpurun_out/bench_n_$1.json
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_n_mag_$1.json 2> gpurun_out/bench_n_mag_$1.err; show gpurun_out/bench_n_mag_$1.json
}
run s3
timeout 300 python bench.py $B --variant 4 > gpurun_out/bench_n_v4.json 2> gpurun_out/bench_n_v4.err; show gpurun_out/bench_n_v4.json
for v in 5 4; do
  FE_TET_VARIANT=$v timeout 200 python scripts/bench_tet.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tet variant $v', d['assembly_ms'], d['assembly_roofline_frac'], 'pcg', d.get('pcg_ms_per_iter'))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02n_s16m_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_n_full.log 2>&1; echo "ncu asm rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tet_assemble_staged' -s 2 -c 1 \
  -o gpurun_out/prof_r02n_tet -f python scripts/bench_tet.py > gpurun_out/ncu_n_tet.log 2>&1; echo "ncu tet rc=$?"
for ns in 4 2; do
  ( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_STAGES=$ns" > /dev/null 2>&1 ); echo "FE_FAN_STAGES=$ns"
  run s$ns
done
( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_STAGES=3 -DFE_FAN_MINB=4" > /dev/null 2>&1 ); echo "stages 3 minb 4"
run s3b4
( cd finite_elements_b200/csrc && touch assemble.cu && make > /dev/null 2>&1 )
