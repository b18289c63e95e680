#!/bin/bash
# Is the A/B difference the library or the position in the sequence (clocks / power / temperature)?
# Runs HEAD and an old build alternately and logs the clocks beside the numbers.
o=gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv -lms 100 > $o/ab_order_clocks.csv &
SMI=$!
for round in 1 2 3; do
for lib in olap_in_memory_b200/libolapgpu.so build/ab/a56024b/libolapgpu.so; do
  tag=$(basename $(dirname $lib))
  date +%T.%N
  OLAP_LIB=$PWD/$lib python bench_ops.py --only "drillup/time-outer day->month sum zero [,drillup/time-outer day->year,drillup/time-outer day->month first zero" --reps 7 2>&1 | grep -o "\"op\": \"[^\"]*\"\|\"ms\": [0-9.]*\|\"frac\": [0-9.]*" | tr "\n" " " | sed "s/\"op\"/\n$tag \"op\"/g"; echo
done
done
kill $SMI
python - <<PY
import csv
rows=list(csv.reader(open("$o/ab_order_clocks.csv")))[1:]
busy=[r for r in rows if float(r[3].split()[0])>300]
print("samples", len(rows), "under load", len(busy))
for r in busy[:: max(1,len(busy)//25)]: print(r)
PY
