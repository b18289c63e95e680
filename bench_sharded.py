#!/usr/bin/env python
"""Sharded-cube measurements (developer tool; run under torchrun, one rank per GPU):
a univac-style cube of 10-item generic dimensions, rows of (dim0, dim1) split across ranks.
  * drillUp of an INNER dimension: shard-local, no collective (weak/strong scaling of HBM)
  * drillUp of dim0 (sharded): partial rollup + all_to_all over NVLink + ordered combine
Times are CUDA events on the rank's stream, max over ranks."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndims", type=int, default=9)  # 10^9 cells
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="run only the operations whose label contains this text")
    ap.add_argument("--exchange", default="", help="comma-separated exchange modes of the sharded rollups to time "
                    "(pull, push, nccl; default: the library's own default)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from olap_in_memory_b200 import GenericDimension, _native, interop
    from olap_in_memory_b200.sharded import ShardedCube

    _native.init(local)
    interop.use_torch_stream()
    dims = []
    for k in range(args.ndims):
        d = GenericDimension(f"dim{k}", "root", [str(i) for i in range(10)])
        d.addAttribute("root", "parity", lambda item: "even" if int(item) % 2 == 0 else "odd")
        dims.append(d)
    cube = ShardedCube(dims, prefix=2)
    for name, rule in (("m_sum", "sum"), ("m_avg", "average"), ("m_first", "first")):
        cube.createStoredMeasure(name, {d.id: rule for d in dims}, "float32", 0)
        v = interop.values_tensor(cube.storedMeasures[name])
        v.uniform_(1.0, 1000.0)
        st = interop.status_tensor(cube.storedMeasures[name])
        if st is not None:
            st.fill_(2)
    torch.cuda.synchronize()
    n_total = cube.storeSize
    measures = len(cube.storedMeasures)

    def timed(fn):
        fn()
        ms = []
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
            del out
        t = torch.tensor([float(np.median(ms))], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rows = []
    last = f"dim{args.ndims - 1}"
    mid = f"dim{args.ndims // 2}"
    from olap_in_memory_b200 import sharded as _sh

    def emit(label, fn, n_out):
        ms = timed(fn)
        row = {"op": label, "n_gpus": world, "cells_in": n_total, "measures": measures, "ms": round(ms, 3),
               "measure_cells_per_s": measures * n_total / (ms * 1e-3),
               "hbm_GBs_per_gpu": round(5 * measures * (n_total + n_out) / world / (ms * 1e-3) / 1e9, 1)}
        if "sharded" in label and world > 1:
            # average travels as (sum, count): 4 planes for 3 measures
            planes = 4
            # every rank holds a partial of the FULL output and sends (W-1)/W of it
            row["exchange"] = _sh.EXCHANGE
            if _sh.EXCHANGE == "pull":  # every rank reads (W-1)/W of the children of its output rows, 3 measures x 5 B
                row["nvlink_bytes_per_gpu"] = 5 * measures * (n_total // world) * (world - 1) // world
            else:
                row["nvlink_bytes_per_gpu"] = 5 * planes * n_out * (world - 1) // world
            row["nvlink_GBs_per_gpu"] = round(row["nvlink_bytes_per_gpu"] / (ms * 1e-3) / 1e9, 1)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)

    for label, fn, n_out in (
        (f"inner {last}->all (shard-local)", lambda: cube.drillUp(last, "all"), n_total // 10),
        (f"mid {mid}->parity (shard-local)", lambda: cube.drillUp(mid, "parity"), n_total // 5),
        ("sharded dim0->all (exchange + combine)", lambda: cube.drillUp("dim0", "all"), n_total // 10),
        ("sharded dim0->parity (exchange + combine)", lambda: cube.drillUp("dim0", "parity"), n_total // 5),
    ):
        if args.only and args.only not in label:
            continue
        modes = [None]
        if "sharded" in label and world > 1 and args.exchange:
            modes = args.exchange.split(",")
        for mode in modes:
            if mode is not None:
                _sh.EXCHANGE = mode
            emit(label, fn, n_out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
