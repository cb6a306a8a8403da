#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list and full captures.
# Usage (under gpurun): bash scripts/gpu_check.sh [tests|bench|tune|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
: > gpurun_out/summary.txt
if [[ $what == all || $what == tests ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
  tail -25 gpurun_out/pytest_gpu.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
  timeout 300 python bench.py --steps 3 --warmup 3 --nx 1024 --ny 512 --no-cpu-baseline > gpurun_out/bench_s1m.json 2>> gpurun_out/bench.err
fi
if [[ $what == all || $what == tune ]]; then
  for v in 2 3; do
    timeout 300 python bench.py --steps 3 --warmup 3 --variant $v --full-solve 0 --no-cpu-baseline --pcg-iters 5 > gpurun_out/bench_v$v.json 2>> gpurun_out/bench.err
  done
  for lib in finite_elements_b200/libfe_b200_*.so; do
    [[ -f $lib ]] || continue
    tag=$(basename $lib .so | sed 's/libfe_b200_//')
    FE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --variant 3 --full-solve 0 --no-cpu-baseline --pcg-iters 5 > gpurun_out/bench_$tag.json 2>> gpurun_out/bench.err
  done
  FE_B200_NO_GRAPH=1 timeout 300 python bench.py --steps 3 --warmup 3 --nx 1024 --ny 512 --full-solve 0 --no-cpu-baseline > gpurun_out/bench_s1m_nograph.json 2>> gpurun_out/bench.err
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --full-solve 0 --no-cpu-baseline --pcg-iters 10 > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assemble|k_spmv|k_pcg_update|k_pcg_pupdate' -s 8 -c 8 \
    -o gpurun_out/prof -f python bench.py --steps 1 --warmup 3 --full-solve 0 --no-cpu-baseline --pcg-iters 3 > gpurun_out/ncu_full.log 2>&1
  echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
