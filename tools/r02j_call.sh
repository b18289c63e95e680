#!/bin/bash
# round 2, call j: lanes kernel (list-driven) — tests, then A/B rows
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_r02j.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02j.log
ONLY="drillup/long,drillup/derived-status customers"
timeout 300 python bench_ops.py --only "$ONLY" --out gpurun_out/ops_r02j_lanes.json > gpurun_out/ops_r02j_lanes.log 2>&1; echo "bench rc=$?"
for ss in 4 8 12 16 24; do
OLAP_LANES_SS=$ss timeout 300 python bench_ops.py --only "drillup/derived-status customers" > gpurun_out/ops_r02j_ss$ss.log 2>&1
done
grep -h '"op"' gpurun_out/ops_r02j_*.log | cut -c1-200
