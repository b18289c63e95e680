#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_r02o.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02o.log
timeout 300 python bench_ops.py --only "dice" --out gpurun_out/ops_r02o_dice.json > gpurun_out/ops_r02o_dice.log 2>&1; echo "bench rc=$?"
OLAP_FLAT=0 timeout 300 python bench_ops.py --only "dice/inner" > gpurun_out/ops_r02o_dice_noflat.log 2>&1
grep -h '"op"' gpurun_out/ops_r02o_dice*.log | cut -c1-220
