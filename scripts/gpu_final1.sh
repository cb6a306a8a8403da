#!/bin/bash
# 1-GPU evidence run: smoke, GPU tests, default bench (+ reference arm), S1M bench, ncu launch list + full captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --nx 1024 --ny 512 --no-cpu-baseline --modal 0 > gpurun_out/bench_s1m.json 2>> gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench_s1m.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --full-solve 0 --no-cpu-baseline --pcg-iters 10 --modal 0 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assemble|k_spmv|k_pcg_update|k_pcg_pupdate|k_bc_apply' -s 8 -c 8 \
  -o gpurun_out/prof -f python bench.py --steps 1 --warmup 3 --full-solve 0 --no-cpu-baseline --pcg-iters 3 --modal 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmm" -s 60 -c 6 \
  -o gpurun_out/prof_spmm -f python bench.py --steps 1 --warmup 3 --full-solve 0 --no-cpu-baseline --pcg-iters 2 --modal 10 > gpurun_out/ncu_spmm.log 2>&1; echo "ncu spmm rc=$?"
