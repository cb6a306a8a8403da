#!/bin/bash
# round 2, call C: phase breakdown of the persistent PCG kernel (CTA 0 timers) + S1M parity test
set -u
mkdir -p gpurun_out
export FE_B200_PERSIST_PROF=1
timeout 300 python bench.py --steps 2 --warmup 3 --full-solve 0 --no-cpu-baseline --modal 0 > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err; tail -3 gpurun_out/bench_prof.err
python scripts/show_bench.py gpurun_out/bench_prof.json
timeout 300 python bench.py --steps 2 --warmup 3 --nx 1024 --ny 512 --full-solve 0 --no-cpu-baseline --modal 0 > gpurun_out/bench_prof_s1m.json 2> gpurun_out/bench_prof_s1m.err; tail -3 gpurun_out/bench_prof_s1m.err
python scripts/show_bench.py gpurun_out/bench_prof_s1m.json
unset FE_B200_PERSIST_PROF
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "s1m_values" > gpurun_out/pytest_s1m.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_s1m.log
