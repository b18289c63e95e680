#!/bin/bash
set -u
mkdir -p gpurun_out
for geo in 0 1; do
OLAP_LANES_GEO=$geo OLAP_LANES_LOADED=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "lanes or long_rows" > gpurun_out/tests_r02n_geo$geo.log 2>&1; echo "tests geo=$geo rc=$?"; tail -1 gpurun_out/tests_r02n_geo$geo.log
ONLY="drillup/long customers,drillup/derived-status customers"
OLAP_LANES_GEO=$geo OLAP_LANES_LOADED=1 timeout 300 python bench_ops.py --only "$ONLY" --out gpurun_out/ops_r02n_geo$geo.json > gpurun_out/ops_r02n_geo$geo.log 2>&1; echo "bench rc=$?"
grep -h '"op"' gpurun_out/ops_r02n_geo$geo.log | cut -c1-160
done
OLAP_LANES_GEO=1 bash tools/ncu_summary.sh lanes_r02n drillup_lanes 1 -- python tools/one_lanes.py sum derived
cat gpurun_out/plain_lanes_r02n.log
