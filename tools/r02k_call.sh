#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lanes or pageable" > gpurun_out/tests_r02k.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02k.log
bash tools/ncu_summary.sh lanes_r02k drillup_lanes 1 -- python tools/one_lanes.py sum derived
cat gpurun_out/plain_lanes_r02k.log
