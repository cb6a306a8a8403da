#!/bin/bash
# quick tuning loop: parity of the assembly variants + bench of variant 3 + ncu of the hot kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pattern or ordered or mid_size or shuffled or edge or empty or bowtie or solution" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python scripts/show_bench.py gpurun_out/bench_quick.json
for lib in finite_elements_b200/libfe_b200_*.so; do
  [[ -f $lib ]] || continue
  tag=$(basename $lib .so | sed 's/libfe_b200_//')
  FE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline > gpurun_out/bench_$tag.json 2>> gpurun_out/bench_quick.err
  python scripts/show_bench.py gpurun_out/bench_$tag.json
done
if [[ ${1:-} == ncu ]]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble|k_spmv' -s 6 -c 3 \
    -o gpurun_out/prof_quick -f python bench.py --steps 1 --warmup 3 --full-solve 0 --no-cpu-baseline --pcg-iters 2 > gpurun_out/ncu_quick.log 2>&1
  echo "ncu rc=$?"
fi
FE_B200_NO_STREAM=1 timeout 300 python bench.py --steps 5 --warmup 3 --full-solve 0 --no-cpu-baseline > gpurun_out/bench_nostream.json 2>> gpurun_out/bench_quick.err
python scripts/show_bench.py gpurun_out/bench_nostream.json
timeout 300 python bench.py --steps 5 --warmup 3 --nx 1024 --ny 512 --full-solve 1 --no-cpu-baseline > gpurun_out/bench_s1m.json 2>> gpurun_out/bench_quick.err
python scripts/show_bench.py gpurun_out/bench_s1m.json
