#!/bin/bash
# round 2, call AM: final validation of the committed build -- smoke, whole GPU suite, default bench line, reference arm
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_am.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_am.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_am.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_am.log
( time timeout 900 python bench.py > gpurun_out/bench_am_full.json 2> gpurun_out/bench_am_full.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_am_full.err
python scripts/show_bench.py gpurun_out/bench_am_full.json 2>/dev/null | head -6
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_am_ref.json 2> gpurun_out/bench_am_ref.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_am_ref.json
