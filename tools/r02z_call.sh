#!/bin/bash
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/tests_r02z.log 2>&1; cat gpurun_out/tests_r02z.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
