#!/bin/bash
# round 2, call B: persistent PCG kernel -- parity tests, then bench with and without it
set -u
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for mode in persist legacy; do
  if [[ $mode == legacy ]]; then export FE_B200_NO_PERSIST=1; else unset FE_B200_NO_PERSIST; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --full-solve ${FULL:-1} --no-cpu-baseline --modal 0 > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench $mode rc=$?"; tail -2 gpurun_out/bench_$mode.err
  python scripts/show_bench.py gpurun_out/bench_$mode.json
  timeout 300 python bench.py --steps 3 --warmup 3 --nx 1024 --ny 512 --full-solve 1 --no-cpu-baseline --modal 0 > gpurun_out/bench_s1m_$mode.json 2>> gpurun_out/bench_$mode.err
  python scripts/show_bench.py gpurun_out/bench_s1m_$mode.json
done
