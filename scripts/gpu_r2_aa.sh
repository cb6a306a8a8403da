#!/bin/bash
# round 2, call AA: 3x3-block streamed SpMV for the tetrahedral PCG
set -u
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_aa.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitize_aa.log
timeout 900 python -m pytest tests/test_tet.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/pytest_aa.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_aa.log
for nb in 0 1; do
  if [[ $nb == 1 ]]; then export FE_B200_NO_BLOCK3=1; else unset FE_B200_NO_BLOCK3; fi
  timeout 200 python scripts/bench_tet.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NO_BLOCK3=$nb tet asm', d['assembly_ms'], d['assembly_roofline_frac'], 'pcg ms/it', d.get('pcg_ms_per_iter'), d.get('pcg_roofline_frac'))"
done
unset FE_B200_NO_BLOCK3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_stream3' -s 20 -c 1 \
  -o gpurun_out/prof_r02aa_spmv3 -f python scripts/bench_tet.py > gpurun_out/ncu_aa.log 2>&1; echo "ncu rc=$?"
