#!/bin/bash
# round 2, call R: rw2 + L1 prefetch of the next chunk first loads
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_r.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitize_q.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_q.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "asm ms", round(d["assembly"]["ms"],4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], "pcg ms/it", round(d["pcg"]["ms_per_iter"],4))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
export FE_B200_FAN_DESIGN=rw
for v in 0 4; do
  timeout 300 python bench.py $B --variant $v > gpurun_out/bench_r_v$v.json 2> gpurun_out/bench_r_v$v.err; show gpurun_out/bench_r_v$v.json
  timeout 300 python bench.py $B --variant $v --kind magnetic > gpurun_out/bench_r_mag_v$v.json 2> gpurun_out/bench_r_mag_v$v.err; show gpurun_out/bench_r_mag_v$v.json
done
timeout 300 python bench.py $B --nx 1024 --ny 512 > gpurun_out/bench_r_s1m.json 2> gpurun_out/bench_r_s1m.err; show gpurun_out/bench_r_s1m.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02r_s16m_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_r_full.log 2>&1; echo "ncu asm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02r_s16m_asm8 -f python bench.py $B --steps 1 --warmup 3 --variant 4 > gpurun_out/ncu_r_full8.log 2>&1; echo "ncu asm8 rc=$?"
