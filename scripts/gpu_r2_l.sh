#!/bin/bash
# round 2, call L: compact (4-byte) fan records -- parity, A/B timing, ncu --set full of the headline step
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_l.log 2>&1; echo "pytest parity rc=$?"; tail -4 gpurun_out/pytest_l.log
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "asm ms", round(d["assembly"]["ms"],4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], "pcg ms/it", round(d["pcg"]["ms_per_iter"],4))
PY
}
for v in 0 4; do
  timeout 300 python bench.py $B --variant $v > gpurun_out/bench_l_v$v.json 2> gpurun_out/bench_l_v$v.err; show gpurun_out/bench_l_v$v.json
  timeout 300 python bench.py $B --variant $v --kind magnetic > gpurun_out/bench_l_mag_v$v.json 2> gpurun_out/bench_l_mag_v$v.err; show gpurun_out/bench_l_mag_v$v.json
done
# occupancy of the scalar instance
for mb in 5 8; do
  ( cd finite_elements_b200/csrc && touch assemble.cu && make EXTRA="-DFE_FAN_MINB_SCALAR=$mb" > /dev/null 2>&1 )
  timeout 300 python bench.py $B --kind magnetic > gpurun_out/bench_l_mag_mb$mb.json 2> gpurun_out/bench_l_mag_mb$mb.err; echo "MINB_SCALAR=$mb"; show gpurun_out/bench_l_mag_mb$mb.json
done
( cd finite_elements_b200/csrc && touch assemble.cu && make > /dev/null 2>&1 )
# ncu: launch list of the bench step, then the full set of the hot kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py $B --steps 2 --warmup 3 > gpurun_out/ncu_l_launch.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02_s16m_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_l_full.log 2>&1; echo "ncu asm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_stream|k_pcg_update|k_pcg_pupdate' -s 60 -c 3 \
  -o gpurun_out/prof_r02_s16m_pcg -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_l_pcg.log 2>&1; echo "ncu pcg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02_mag -f python bench.py $B --steps 1 --warmup 3 --kind magnetic > gpurun_out/ncu_l_mag.log 2>&1; echo "ncu mag rc=$?"
ls -la gpurun_out/*.ncu-rep
