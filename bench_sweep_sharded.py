#!/usr/bin/env python
"""Config 5 of BASELINE.json beyond one GPU (developer tool; run under torchrun, one rank per GPU):
8 stored measures with all aggregation types + 4 computed measures on a cube
[g_outer x64 (sharded rows), time day x3652, g_inner], sizes up to 1e10 cells: drillUp time day -> month
for the 8 measures in ONE batched shard-local call, then the 4 formulas (one of them needs a
whole-cube total: one all-reduce).  Times are CUDA events on the rank's stream, max over ranks."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METHODS = ["sum", "average", "highest", "lowest", "first", "last", "product", "sum"]
FORMULAS = ["(m0 + m7) / m1", "m2 - m3", "m4 || m5", "isNaN(m6) + m0 / m0__total"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1e9,1e10")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from olap_in_memory_b200 import GenericDimension, TimeDimension, _native, interop
    from olap_in_memory_b200.sharded import ShardedCube

    _native.init(local)
    lib = _native.lib()
    interop.use_torch_stream()
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for target in [float(x) for x in args.sizes.split(",")]:
        g_inner = max(4, int(round(target / 3652 / 64 / 4)) * 4)
        n_total = 64 * 3652 * g_inner
        per_gpu = 5 * len(METHODS) * n_total / world * 1.12  # inputs + rolled outputs + formula results
        if per_gpu > 150e9:
            if rank == 0:
                print(json.dumps({"cells": n_total, "n_gpus": world, "skipped": f"needs {per_gpu / 1e9:.0f} GB per GPU"}), flush=True)
            continue
        dims = [GenericDimension("g_outer", "root", [str(i) for i in range(64)]), TimeDimension("time", "day", "2010-01-01", "2019-12-31"),
                GenericDimension("g_inner", "root", [str(i) for i in range(g_inner)])]
        cube = ShardedCube(dims, prefix=1)
        for k, method in enumerate(METHODS):
            cube.createStoredMeasure(f"m{k}", {"time": method}, "float32", 0)
            v = interop.values_tensor(cube.storedMeasures[f"m{k}"])
            v.uniform_(0.9, 1.1) if method == "product" else v.uniform_(1.0, 1000.0)
            interop.status_tensor(cube.storedMeasures[f"m{k}"]).fill_(2)
        for k, f in enumerate(FORMULAS):
            cube.createComputedMeasure(f"c{k}_f", f)
        torch.cuda.synchronize()

        def step():
            rolled = cube.drillUp("time", "month")
            ms = lib.olap_last_op_ms()
            step.path = lib.olap_last_op_path().decode()
            outs = [rolled.getLocalStore(f"c{k}_f") for k in range(len(FORMULAS))]
            return ms, outs

        step()
        k_ms, w_ms = [], []
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ms, outs = step()
            e1.record()
            torch.cuda.synchronize()
            k_ms.append(ms)
            w_ms.append(e0.elapsed_time(e1))
            del outs
        t = torch.tensor([float(np.median(k_ms)), float(np.median(w_ms))], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k, w = float(t[0].item()), float(t[1].item())
        n_local = cube.localSize
        algo = 5 * len(METHODS) * (n_local + n_local // 3652 * 120)
        if rank == 0:
            print(json.dumps({"cells": n_total, "n_gpus": world, "shape": [64, 3652, g_inner], "measures": len(METHODS), "computed": len(FORMULAS),
                              "path": step.path, "drillup_kernel_ms": round(k, 3), "step_ms": round(w, 3),
                              "measure_cells_per_s_kernel": len(METHODS) * n_total / (k * 1e-3),
                              "measure_cells_per_s_step": len(METHODS) * n_total / (w * 1e-3),
                              "hbm_GBs_per_gpu": round(algo / (k * 1e-3) / 1e9, 1), "frac_per_gpu": round(algo / (k * 1e-3) / 1e9 / peak, 3)}), flush=True)
        del cube
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
