#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "not lanes and not fullsize" > gpurun_out/tests_r02zz.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests_r02zz.log; grep -n "^E " gpurun_out/tests_r02zz.log | head -5
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
