#!/bin/bash
# round 2, call i: lanes kernel + pageable ring — tests, then A/B rows
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_r02i.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02i.log
ONLY="drillup/long,drillup/derived-status customers,load,sparse,total,boundary"
timeout 300 python bench_ops.py --only "$ONLY" --out gpurun_out/ops_r02i_lanes.json > gpurun_out/ops_r02i_lanes.log 2>&1; echo "bench rc=$?"
OLAP_LANES=0 OLAP_HOST_PIPE=0 timeout 300 python bench_ops.py --only "$ONLY" --out gpurun_out/ops_r02i_nolanes.json > gpurun_out/ops_r02i_nolanes.log 2>&1; echo "bench0 rc=$?"
for t in 2 4 8 16; do
  OLAP_COPY_THREADS=$t timeout 200 python bench_ops.py --only "boundary" --out gpurun_out/ops_r02i_threads$t.json > gpurun_out/ops_r02i_threads$t.log 2>&1
done
nproc; grep -h '"op"' gpurun_out/ops_r02i_*.log | cut -c1-330
