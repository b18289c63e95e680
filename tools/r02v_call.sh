#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not lanes" > gpurun_out/tests_r02v.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02v.log
timeout 300 python bench_ops.py --only "reorder" --out gpurun_out/ops_r02v.json > gpurun_out/ops_r02v.log 2>&1; echo "bench rc=$?"
grep -h '"op"' gpurun_out/ops_r02v.log | cut -c1-180
