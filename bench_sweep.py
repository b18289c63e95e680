#!/usr/bin/env python
"""Config 5 of BASELINE.json (developer tool): multi-measure rollup sweep — 8 stored
measures with all aggregation types (sum, average, highest, lowest, first, last, product,
sum) + 4 computed measures, cube [time day x3652, g], sizes 1e6 .. 1e9 cells on one B200,
drillUp time day->month in ONE batched call, computed measures evaluated on the result.
The CPU column is the C port of the reference algorithm (1 thread) at sizes it finishes
quickly.   python bench_sweep.py [--max 1e9] [--cpu-max 2e7]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METHODS = ["sum", "average", "highest", "lowest", "first", "last", "product", "sum"]
FORMULAS = ["(m0 + m7) / m1", "m2 - m3", "m4 || m5", "isNaN(m6) + m0 / m0__total"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max", type=float, default=1e9)
    ap.add_argument("--cpu-max", type=float, default=2e7)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    from olap_in_memory_b200 import Cube, GenericDimension, TimeDimension, _native, interop
    from oracle.c_oracle import COracleStore

    _native.init(0)
    lib = _native.lib()
    interop.use_torch_stream()
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []
    # generic axis of a multiple of 4 items (128-bit accesses); at 1e8 cells also an odd one
    # (the scalar variant of the same kernel), so that both are on record
    shapes = []
    for target in (1e6, 1e7, 1e8, 1e9, 1e10):
        if target > args.max:
            break
        g4 = max(4, int(round(target / 3652 / 4)) * 4)
        shapes.append(g4)
        if target == 1e8:
            shapes.append(g4 + 1)
    for g in shapes:
        dims = [TimeDimension("time", "day", "2010-01-01", "2019-12-31"),
                GenericDimension("g", "root", [str(i) for i in range(g)])]
        cube = Cube(dims)
        for k, method in enumerate(METHODS):
            cube.createStoredMeasure(f"m{k}", {"time": method}, "float32", 0)
            v = interop.values_tensor(cube.storedMeasures[f"m{k}"])
            if method == "product":
                v.uniform_(0.9, 1.1)
            else:
                v.uniform_(1.0, 1000.0)
            st = interop.status_tensor(cube.storedMeasures[f"m{k}"])
            if st is not None:
                st.fill_(2)
        for k, f in enumerate(FORMULAS):
            cube.createComputedMeasure(f"c{k}_f", f)
        torch.cuda.synchronize()
        n = cube.storeSize

        def step():
            rolled = cube.drillUp("time", "month")
            ms = lib.olap_last_op_ms()
            step.path = lib.olap_last_op_path().decode()
            outs = [rolled.evaluateToStore(f"c{k}_f") for k in range(len(FORMULAS))]
            return ms, outs

        step()
        kernel_ms, wall = [], []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ms, outs = step()
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            kernel_ms.append(ms)
            del outs
        k_ms, w_ms = float(np.median(kernel_ms)), float(np.median(wall))
        algo = 5 * len(METHODS) * (n + 120 * g)
        row = {"cells": n, "shape": [3652, g], "inner_run": "multiple of 4" if g % 4 == 0 else ("even" if g % 2 == 0 else "odd"),
               "path": getattr(step, "path", ""), "measures": len(METHODS), "computed": len(FORMULAS), "drillup_kernel_ms": round(k_ms, 4),
               "step_wall_ms": round(w_ms, 3), "measure_cells_per_s_kernel": len(METHODS) * n / (k_ms * 1e-3),
               "measure_cells_per_s_step": len(METHODS) * n / (w_ms * 1e-3),
               "GBs": round(algo / (k_ms * 1e-3) / 1e9, 1), "frac": round(algo / (k_ms * 1e-3) / 1e9 / peak, 3)}
        if n <= args.cpu_max:
            month = np.asarray(dims[0].getGroupIndexFromRootIndexMap("month"), np.int32)
            ident = np.arange(g, dtype=np.int32)
            stores = []
            for k, method in enumerate(METHODS):
                s = COracleStore(n, "float32", 0.0)
                s.set_data_f32(cube.storedMeasures[f"m{k}"].data_f32())
                stores.append(s)
            t0 = time.perf_counter()
            for s, method in zip(stores, METHODS):
                s.drillUp_lowered([3652, g], [120, g], [month, ident], method)
            cpu_s = time.perf_counter() - t0
            row["cpu_port_1thread_cells_per_s"] = len(METHODS) * n / cpu_s
            del stores
        else:
            row["cpu_port_1thread_cells_per_s"] = None  # beyond the reference's 2^24-entry Map
        rows.append(row)
        print(json.dumps(row), flush=True)
        del cube
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
