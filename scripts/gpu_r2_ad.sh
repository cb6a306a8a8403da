#!/bin/bash
# round 2, call AD: final build incl. 3x3 SpMV and node-blocked SpMM -- whole GPU suite, full bench line, ncu launch list + full sets (traffic source)
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_ad.log 2>&1; echo "sanitizer rc=$?"; tail -2 gpurun_out/sanitize_x.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ad.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ad.log
( time timeout 900 python bench.py > gpurun_out/bench_ad_full.json 2> gpurun_out/bench_ad_full.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_ad_full.err
python scripts/show_bench.py gpurun_out/bench_ad_full.json 2>/dev/null | head -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ad_launches.csv \
  python bench.py $B --steps 2 --warmup 3 > gpurun_out/ncu_ad_launch.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02ad_s16m_asm -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_ad_asm.log 2>&1; echo "ncu asm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_stream|k_pcg_update|k_pcg_pupdate' -s 60 -c 3 \
  -o gpurun_out/prof_r02ad_s16m_pcg -f python bench.py $B --steps 1 --warmup 3 > gpurun_out/ncu_ad_pcg.log 2>&1; echo "ncu pcg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_fan' -s 3 -c 1 \
  -o gpurun_out/prof_r02ad_mag -f python bench.py $B --steps 1 --warmup 3 --kind magnetic > gpurun_out/ncu_ad_mag.log 2>&1; echo "ncu mag rc=$?"
ls -la gpurun_out/prof_r02ad*.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tet_assemble_pipe' -s 2 -c 1 \
  -o gpurun_out/prof_r02ad_tet -f python scripts/bench_tet.py > gpurun_out/ncu_ad_tet.log 2>&1; echo "ncu tet rc=$?"
