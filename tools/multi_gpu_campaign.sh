#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus G): hardware verification of the sharded path, then the
# driver-style bench at every N in "$@" (default: all powers of two up to the visible GPU count).
# Outputs land in gpurun_out/ (tag = $TAG, default r02).
set -u
tag=${TAG:-r02}
o=gpurun_out
G=$(nvidia-smi -L | wc -l)
ns=${@:-$(for n in 2 4 8; do [ $n -le $G ] && echo $n; done)}
free -g | head -2 > $o/host_$tag.txt; nproc >> $o/host_$tag.txt; nvidia-smi topo -m >> $o/host_$tag.txt 2>&1
port=29600
run() { # run N cmd...
  local n=$1; shift; port=$((port + 1))
  timeout ${TMO:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"
}
for n in $ns; do
  for mode in ${CHECK_MODES-pull pull2 push nccl}; do
    OLAP_SHARDED_EXCHANGE=$mode run $n tests/gpu_sharded_check.py > $o/shard_check_${tag}_n${n}_$mode.log 2>&1
    echo "check n=$n $mode: $(grep gpu_sharded_check $o/shard_check_${tag}_n${n}_$mode.log | tail -1)"
  done
done
for n in $ns; do
  [ -n "${NOBENCH:-}" ] && continue
  if [ -n "${SMALL:-}" ]; then
    OLAP_BENCH_NDIMS=$SMALL run $n bench.py --gpus $n --steps 5 --warmup 3 > $o/bench_${tag}_n${n}_small.json 2> $o/bench_${tag}_n${n}_small.err
    echo "bench small n=$n: $(cut -c1-400 $o/bench_${tag}_n${n}_small.json)"; tail -3 $o/bench_${tag}_n${n}_small.err
  fi
  run $n bench.py --gpus $n --steps ${STEPS:-20} --warmup 5 > $o/bench_${tag}_n$n.json 2> $o/bench_${tag}_n$n.err
  echo "bench n=$n: $(cat $o/bench_${tag}_n$n.json | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d["sharded"]
    print(d["value"], "ms/step", d["ms_per_step"], "inner", s["inner_rollup"]["ms"], s["inner_rollup"]["hbm_frac"], "outer", s["sharded_rollup"]["ms"], s["sharded_rollup"]["kernel_ms"], s["sharded_rollup"]["nvlink_GBs_per_gpu"], "guard", s["parity_guard"]["mismatches"], s["parity_guard"]["first_mismatches"], "e2e", d["e2e"]["ms_per_step"])
except Exception as e: print("no line", e)')"
  tail -3 $o/bench_${tag}_n$n.err
  if [ -n "${MODES:-}" ]; then
    run $n bench_sharded.py --ndims ${NDIMS:-10} --only "dim0->all" --exchange $MODES > $o/sharded_${tag}_n$n.jsonl 2> $o/sharded_${tag}_n$n.err
    cat $o/sharded_${tag}_n$n.jsonl; tail -2 $o/sharded_${tag}_n$n.err
  fi
done
