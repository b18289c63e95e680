#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "innermost_axis_alone" > gpurun_out/tests_r02p.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02p.log
bash tools/ncu_summary.sh flat_r02p gather_inner_flat 1 -- python tools/one_dice_inner.py
cat gpurun_out/plain_flat_r02p.log
OLAP_FLAT=0 bash tools/ncu_summary.sh rows_r02p gather_rows 1 -- python tools/one_dice_inner.py
cat gpurun_out/plain_rows_r02p.log
