#!/bin/bash
# Runs on the GPU box (through gpurun): the round's evidence in one call.  Outputs land in gpurun_out/.
set -u
tag=${1:-r01d}
o=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $o/tests_$tag.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/smoke_$tag.log 2>&1
timeout 600 python bench.py --impl reference > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err
timeout 600 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err
timeout 900 python bench_ops.py --out $o/ops_$tag.json > $o/ops_$tag.log 2>&1
timeout 600 python bench_cube_benchmark.py --out $o/cube_benchmark_$tag.json > $o/cube_benchmark_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv python bench.py --steps 3 --warmup 3 > $o/ncu_launches_$tag.log 2>&1
timeout 600 tools/ncu_summary.sh pair_$tag transpose_pair 1 -- python bench_ops.py --only reorder/reverse --reps 1
timeout 600 tools/ncu_summary.sh long_$tag drillup_long_kernel 1 -- python bench_ops.py --only "drillup/long collapse sum" --reps 1
tail -2 $o/tests_$tag.log; cat $o/smoke_$tag.log | tail -1; cat $o/bench_$tag.json; cat $o/bench_ref_$tag.json; tail -3 $o/cube_benchmark_$tag.log
