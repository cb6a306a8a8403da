#!/bin/bash
# Under `gpurun --gpus N`: distributed parity worker + N-GPU bench lines, for both transports
# (default: peer-memory halo + fused all-reduce; FE_B200_NO_P2P=1: NCCL send/recv + all-reduce).
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout ${T:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
for mode in ${MODES:-p2p nccl}; do
  if [[ $mode == nccl ]]; then export FE_B200_NO_P2P=1; else unset FE_B200_NO_P2P; fi
  run tests/dist_gpu_worker.py 96 64 > gpurun_out/dist_worker_${N}_$mode.log 2>&1; echo "[$mode] dist worker rc=$?"; grep -E "DIST-OK|Error|error" gpurun_out/dist_worker_${N}_$mode.log | head -3
  run tests/dist_gpu_worker.py 301 77 >> gpurun_out/dist_worker_${N}_$mode.log 2>&1; echo "[$mode] dist worker(301x77) rc=$?"; grep -E "DIST-OK" gpurun_out/dist_worker_${N}_$mode.log | tail -1
  run tests/dist_gpu_worker.py 301 177 mag >> gpurun_out/dist_worker_${N}_$mode.log 2>&1; echo "[$mode] dist worker(magnetic 301x177) rc=$?"; grep -E "DIST-OK" gpurun_out/dist_worker_${N}_$mode.log | tail -1
  run tests/dist_gpu_worker.py 24 12 tet >> gpurun_out/dist_worker_${N}_$mode.log 2>&1; echo "[$mode] dist worker(tetrahedra 24x12x4) rc=$?"; grep -E "DIST-OK" gpurun_out/dist_worker_${N}_$mode.log | tail -1
  if [[ $mode == p2p ]]; then  # the non-fused exchange kernel (k_halo_ll) behind the fallback SpMV kernels
    FE_B200_NO_STREAM=1 run tests/dist_gpu_worker.py 96 64 > gpurun_out/dist_worker_${N}_nostream.log 2>&1; echo "[p2p, no stream] dist worker rc=$?"; grep -E "DIST-OK" gpurun_out/dist_worker_${N}_nostream.log | tail -1
    FE_B200_NO_STREAM=1 run tests/dist_gpu_worker.py 96 64 mag >> gpurun_out/dist_worker_${N}_nostream.log 2>&1; echo "[p2p, no stream] dist worker(magnetic) rc=$?"; grep -E "DIST-OK" gpurun_out/dist_worker_${N}_nostream.log | tail -1
  fi
  [[ ${SKIP_BENCH:-0} == 1 ]] && continue
  run bench.py --gpus $N --steps 3 --warmup 3 --nx 1024 --ny 512 --full-solve 1 --no-cpu-baseline > gpurun_out/bench_s1m_g${N}_$mode.json 2> gpurun_out/bench_g${N}_$mode.err; echo "[$mode] bench s1m rc=$?"
  run bench.py --gpus $N --steps 3 --warmup 3 --full-solve ${FULL:-0} --no-cpu-baseline > gpurun_out/bench_g${N}_$mode.json 2>> gpurun_out/bench_g${N}_$mode.err; echo "[$mode] bench s16m rc=$?"; grep -E "Error|error" gpurun_out/bench_g${N}_$mode.err | head -3
  python scripts/show_bench.py gpurun_out/bench_s1m_g${N}_$mode.json gpurun_out/bench_g${N}_$mode.json
done
