#!/bin/bash
# A/B of the single-measure mid/split drillUp rows across commits (VERDICT r01 "weak" item 9):
# build/ab/<sha>/libolapgpu.so are builds of older commits (made by hand with git worktree),
# timed with HEAD's bench_ops.py through OLAP_LIB.  Runs on the GPU box through gpurun.
set -u
o=gpurun_out
for lib in ${AB_LIBS:-build/ab/*/libolapgpu.so} olap_in_memory_b200/libolapgpu.so; do
  tag=$(basename $(dirname $lib))
  OLAP_LIB=$PWD/$lib timeout 300 python bench_ops.py --only drillup/time-outer --reps 7 --out $o/ab_$tag.json > $o/ab_$tag.log 2>&1
  echo "== $tag"; python - <<PY
import json
for r in json.load(open("$o/ab_$tag.json")):
    print(f"{r['frac']:.3f} {r['ms']:.4f} {r['path']:<24} {r['op']}")
PY
done
