#!/bin/bash
# Runs on the GPU box (through gpurun): the round's single-GPU evidence in one call.  Outputs land in gpurun_out/.
set -u
tag=${1:-r02}
o=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $o/tests_$tag.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/smoke_$tag.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err
timeout 600 python bench.py --steps 20 --warmup 5 > $o/bench_$tag.json 2> $o/bench_$tag.err
timeout 900 python bench_ops.py --out $o/ops_$tag.json > $o/ops_$tag.log 2>&1
timeout 600 python bench_sweep.py --max 1e9 --cpu-max 2e7 --out $o/sweep_config5_$tag.json > $o/sweep_config5_$tag.log 2>&1
timeout 600 python bench_config3.py --out $o/config3_$tag.json > $o/config3_$tag.log 2>&1
timeout 600 python bench_cube_benchmark.py --out $o/cube_benchmark_$tag.json > $o/cube_benchmark_$tag.log 2>&1
timeout 300 python tools/membw.py > $o/membw_$tag.json 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > $o/plain_$tag.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu > $o/ncu_launches_$tag.log 2>&1
# full captures are reduced to text / CSV on the box (a .ncu-rep is ~50 MB; gpurun brings back 64 MB at most)
summarise() { ncu -i /tmp/$1.ncu-rep --page details 2>/dev/null | grep -vE '^\s*$' > $o/ncu_details_$1.txt; ncu -i /tmp/$1.ncu-rep --page raw --csv 2>/dev/null > $o/ncu_raw_$1.csv; rm -f /tmp/$1.ncu-rep; }
timeout 600 ncu --set full --clock-control none -k regex:drillup_mid_kernel -s 2 -c 1 -o /tmp/mid_$tag python bench.py --steps 3 --warmup 3 --no-cpu > $o/ncu_mid_$tag.log 2>&1; summarise mid_$tag
timeout 600 ncu --set full --clock-control none -k regex:drilldown_inner_kernel -c 1 -o /tmp/down_$tag python bench_ops.py --only "drilldown/mid month->day float [50000" --reps 1 > $o/ncu_down_$tag.log 2>&1; summarise down_$tag
tail -2 $o/tests_$tag.log; tail -1 $o/smoke_$tag.log; cut -c1-400 $o/bench_$tag.json; cut -c1-300 $o/bench_ref_$tag.json; tail -3 $o/cube_benchmark_$tag.log | cut -c1-300
