#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not lanes" > gpurun_out/tests_r02y.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02y.log; grep -n "^E " gpurun_out/tests_r02y.log | head -8
timeout 300 python bench_ops.py --only "reorder/rotate" --out gpurun_out/ops_r02y.json > gpurun_out/ops_r02y.log 2>&1; echo "bench rc=$?"
OLAP_FLAT=0 timeout 300 python bench_ops.py --only "reorder/rotate front" > gpurun_out/ops_r02y_noflat.log 2>&1
grep -h '"op"' gpurun_out/ops_r02y.log gpurun_out/ops_r02y_noflat.log | cut -c8-170
