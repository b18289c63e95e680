#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not lanes" > gpurun_out/tests_r02u.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02u.log
timeout 300 python bench_ops.py --only "reorder" --out gpurun_out/ops_r02u.json > gpurun_out/ops_r02u.log 2>&1; echo "bench rc=$?"
OLAP_PAIR_ASYNC=0 timeout 300 python bench_ops.py --only "reorder/reverse,reorder/derived-status reverse" > gpurun_out/ops_r02u_sync.log 2>&1
grep -h '"op"' gpurun_out/ops_r02u.log gpurun_out/ops_r02u_sync.log | cut -c1-180
