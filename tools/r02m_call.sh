#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "lanes or pageable or long_rows" > gpurun_out/tests_r02m.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02m.log
ONLY="drillup/long customers,drillup/derived-status customers"
timeout 300 python bench_ops.py --only "$ONLY" --out gpurun_out/ops_r02m_lanes.json > gpurun_out/ops_r02m_lanes.log 2>&1; echo "bench rc=$?"
grep -h '"op"' gpurun_out/ops_r02m_*.log | cut -c1-200
bash tools/ncu_summary.sh lanes_r02m drillup_lanes 1 -- python tools/one_lanes.py sum derived
cat gpurun_out/plain_lanes_r02m.log
