#!/bin/bash
# Final 8-GPU call of round 2: the driver-style bench at N = 8 and N = 4 with the code as committed.
set -u
o=gpurun_out
run() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) "$@"; }
for n in 8 4; do
  run $n bench.py --gpus $n --steps 20 --warmup 5 > $o/bench_r02f_n$n.json 2> $o/bench_r02f_n$n.err
  python - <<PY
import json
d=json.loads(open("$o/bench_r02f_n$n.json").read().strip().splitlines()[-1]); s=d["sharded"]
print("N=$n", d["value"], d["ms_per_step"], "inner", s["inner_rollup"]["ms"], s["inner_rollup"]["hbm_frac"], "outer", s["sharded_rollup"], "alt", s["alternative_exchanges_ms"], "guard", s["parity_guard"]["mismatches"], s["parity_guard"]["first_mismatches"], "e2e", d["e2e"]["ms_per_step"])
PY
  tail -2 $o/bench_r02f_n$n.err
done
