#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tet.py -m gpu -x -q > gpurun_out/pytest_tet.log 2>&1; echo "pytest tet rc=$?"; tail -3 gpurun_out/pytest_tet.log
timeout 200 python scripts/bench_tet.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', d['assembly_ms'], d['assembly_roofline_frac'])"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_tet_assemble_table|k_tet_gradient' -c 2 python scripts/bench_tet.py 2>&1 | grep -E "gpu__time|dram__|l1tex" | head -24
