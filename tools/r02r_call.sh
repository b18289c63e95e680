#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "innermost_axis_alone or test_case_matches_oracle" > gpurun_out/tests_r02r.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02r.log
timeout 300 python bench_ops.py --only "dice/inner" --out gpurun_out/ops_r02r_dice.json > gpurun_out/ops_r02r_dice.log 2>&1; echo "bench rc=$?"
grep -h '"op"' gpurun_out/ops_r02r_dice*.log | cut -c1-220
bash tools/ncu_summary.sh flat_r02r gather_inner_flat 1 -- python tools/one_dice_inner.py
cat gpurun_out/plain_flat_r02r.log
