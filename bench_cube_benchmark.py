#!/usr/bin/env python
"""The reference's own benchmark (test/cube-benchmark.js) through the drop-in: the same
cubes (createLargeTestCube(10, 4, 3[, rate]) = 10 generic dimensions x 4 items = 1 048 576
cells, 3 float32 measures, fill rates 1 / 0.5 / 0.25 / 0.1 — test/helpers/
create-large-test-cube.js), the same eight operations, the same batchRun(10) averaging of
wall-clock milliseconds around the public `Cube` call.

Two columns per line: `gpu_ms` = this repo (GpuStore behind Cube, synchronous mode, on
cuda:0) and `cpu_ms` = the C port of src/store/in-memory.js behind the SAME Cube host
code (oracle/, 1 thread) — the reference itself needs Node.js, which this image lacks.
  python bench_cube_benchmark.py [--size 4] [--dims 10] [--times 10] [--no-cpu] [--out f.json]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def create_large_test_cube(store_cls, n_dims, size, n_measures, rate=1.0, first=0, seed=0):
    """test/helpers/create-large-test-cube.js:3-47 (random cells set to 1)."""
    from olap_in_memory_b200 import Cube, GenericDimension

    dims = [GenericDimension(f"dimension{i}", "root", [f"dimension{i}-item{j}" for j in range(size)]) for i in range(n_dims)]
    cube = Cube(dims, store_cls)
    rng = np.random.default_rng(seed)
    n = cube.storeSize
    data = np.zeros(n, np.float32)
    data[rng.permutation(n)[: int(rate * n)]] = 1.0
    for i in range(first, first + n_measures):
        cube.createStoredMeasure(f"measure{i}", {}, "float32", 0)
        cube.setData(f"measure{i}", data)
    return cube


def batch_run(fn, times):
    """cube-benchmark.js:5-19."""
    fn()  # the reference's first iteration pays JIT warm-up inside the average; ours pays NVRTC/alloc here
    total = 0.0
    for _ in range(times):
        t0 = time.perf_counter()
        fn()
        total += time.perf_counter() - t0
    return total / times * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4)
    ap.add_argument("--dims", type=int, default=10)
    ap.add_argument("--times", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-only", action="store_true", help="time the C port alone (no GPU in reach)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    from olap_in_memory_b200 import GenericDimension
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200.store import GpuStore

    arms = []
    if not args.cpu_only:
        N.init(0)
        arms.append(("gpu_ms", GpuStore))
    if not args.no_cpu:
        from oracle.c_oracle import COracleStore  # the CPU column only; never on the product path

        arms.append(("cpu_ms", COracleStore))
    rates = (1.0, 0.5, 0.25, 0.1)
    last = args.dims - 1
    new_dim = GenericDimension("dimension-new", "root", [f"dimension-new-item{j}" for j in range(5)])
    ops = [
        ("slice whole dimension", lambda c: c.slice("dimension0", "all", "all")),
        ("slice dimension item", lambda c: c.slice("dimension3", "root", "dimension3-item2")),
        ("collapse", lambda c: c.collapse()),
        ("reorder (reversed)", lambda c: c.reorderDimensions(list(reversed(c.dimensionIds)))),
        ("dice 2 items", lambda c: c.dice("dimension2", "root", ["dimension2-item2", "dimension2-item3"])),
        ("addDimension (5 items)", lambda c: c.addDimension(new_dim)),
        ("removeDimension", lambda c: c.removeDimension(f"dimension{min(4, last)}")),
    ]
    rows = []
    cubes = {}
    for col, cls in arms:
        cubes[col] = [create_large_test_cube(cls, args.dims, args.size, 3, r, 3 if r == 0.1 else 0, seed=k)
                      for k, r in enumerate(rates)]
    for name, op in ops:
        for k, r in enumerate(rates):
            row = {"op": name, "fill": r, "cells": cubes[arms[0][0]][k].storeSize, "measures": 3}
            for col, _ in arms:
                row[col] = round(batch_run(lambda: op(cubes[col][k]), args.times), 3)
            if "cpu_ms" in row and "gpu_ms" in row:
                row["cpu_over_gpu"] = round(row["cpu_ms"] / row["gpu_ms"], 1)
            rows.append(row)
            print(json.dumps(row), flush=True)
    row = {"op": "compose sparse10 with sparse50", "cells": cubes[arms[0][0]][0].storeSize}
    for col, _ in arms:
        row[col] = round(batch_run(lambda: cubes[col][3].compose(cubes[col][1]), args.times), 3)
    if "cpu_ms" in row and "gpu_ms" in row:
        row["cpu_over_gpu"] = round(row["cpu_ms"] / row["gpu_ms"], 1)
    rows.append(row)
    print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"bench": "test/cube-benchmark.js", "cpu": "C port of in-memory.js, 1 thread", "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
