#!/bin/bash
# round 2, call AN: ncu --set full of the magnetic PCG iteration (scalar streamed SpMV + vector kernels)
set -u
mkdir -p gpurun_out
B="--full-solve 0 --modal 0 --extras 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_stream1|k_pcg_update|k_pcg_pupdate' -s 60 -c 3 \
  -o gpurun_out/prof_r02an_mag_pcg -f python bench.py $B --steps 1 --warmup 3 --kind magnetic > gpurun_out/ncu_an.log 2>&1; echo "ncu rc=$?"
