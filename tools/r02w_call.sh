#!/bin/bash
set -u
mkdir -p gpurun_out
ONLY="reorder/derived-status swap,reorder/derived-status rotate,reorder/swap inner,reorder/rotate,dice/inner"
for rep in 1 2; do
timeout 300 python bench_ops.py --only "$ONLY" --out gpurun_out/ops_r02w_flat_$rep.json > gpurun_out/ops_r02w_flat_$rep.log 2>&1
OLAP_FLAT=0 timeout 300 python bench_ops.py --only "$ONLY" --out gpurun_out/ops_r02w_noflat_$rep.json > gpurun_out/ops_r02w_noflat_$rep.log 2>&1
done
grep -h '"op"' gpurun_out/ops_r02w_*.log | cut -c8-150
