#!/bin/bash
# round 2, call M: fan kernel with smem-staged neighbour coordinates; staged tetrahedral assembly
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_m.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitize_m.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tet.py -m gpu -x -q > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_m.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "asm ms", round(d["assembly"]["ms"],4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], "pcg ms/it", round(d["pcg"]["ms_per_iter"],4))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
for v in 0 4; do
  timeout 300 python bench.py $B --variant $v > gpurun_out/bench_m_v$v.json 2> gpurun_out/bench_m_v$v.err; show gpurun_out/bench_m_v$v.json
  timeout 300 python bench.py $B --variant $v --kind magnetic > gpurun_out/bench_m_mag_v$v.json 2> gpurun_out/bench_m_mag_v$v.err; show gpurun_out/bench_m_mag_v$v.json
done
for v in 5 4; do
  FE_TET_VARIANT=$v timeout 200 python scripts/bench_tet.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tet variant $v', d['assembly_ms'], d['assembly_roofline_frac'], 'pcg', d.get('pcg_ms_per_iter'))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02m_s16m_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_m_full.log 2>&1; echo "ncu asm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02m_mag -f python bench.py $B --steps 1 --warmup 3 --kind magnetic > gpurun_out/ncu_m_mag.log 2>&1; echo "ncu mag rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tet_assemble_staged' -s 2 -c 1 \
  -o gpurun_out/prof_r02m_tet -f python scripts/bench_tet.py > gpurun_out/ncu_m_tet.log 2>&1; echo "ncu tet rc=$?"
ls -la gpurun_out/prof_r02m*.ncu-rep
