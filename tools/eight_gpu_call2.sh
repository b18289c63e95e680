#!/bin/bash
# Second 8-GPU call of round 2: the driver-style bench at N = 8 with the balanced row split, then
# config 5 beyond one GPU (8 stored + 4 computed measures, up to 1e10 cells) on 8, 4 and 2 GPUs.
set -u
o=gpurun_out
run() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) "$@"; }
run 8 bench.py --gpus 8 --steps 20 --warmup 5 > $o/bench_r02b_n8.json 2> $o/bench_r02b_n8.err
python - <<PY
import json
d=json.loads(open("$o/bench_r02b_n8.json").read().strip().splitlines()[-1]); s=d["sharded"]
print("N=8", d["value"], d["ms_per_step"], "inner", s["inner_rollup"]["ms"], s["inner_rollup"]["hbm_frac"], "outer", s["sharded_rollup"]["ms"], s["sharded_rollup"]["kernel_ms"], s["sharded_rollup"]["nvlink_GBs_per_gpu"], s["rows_per_rank_out"], "alt", s["alternative_exchanges_ms"], "guard", s["parity_guard"]["mismatches"], s["parity_guard"]["first_mismatches"], "e2e", d["e2e"]["ms_per_step"])
PY
tail -2 $o/bench_r02b_n8.err
for n in 8 4 2; do
  run $n bench_sweep_sharded.py --sizes 1e9,1e10 > $o/sweep_sharded_r02_n$n.jsonl 2> $o/sweep_sharded_r02_n$n.err
  cat $o/sweep_sharded_r02_n$n.jsonl; tail -2 $o/sweep_sharded_r02_n$n.err
done
