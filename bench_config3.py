#!/usr/bin/env python
"""Config 3 of BASELINE.json at its stated size (developer tool, one B200): the 6-dimension cube
100^3 x 10^3 (1e9 cells, one float32 measure with its status plane), then
  dice      dimension a to every other item                     1e9 -> 5e8 cells
  reorder   to the reversed axis order                          1e9 cells
  drillDown time month -> day (10 months -> 304 days) of the DICED cube   5e8 -> 1.52e10 cells (76 GB)
Kernel time from the library's CUDA-event bracket, algorithmic bytes per SURVEY.md §8d, fraction of the
measured HBM copy bandwidth.  Full-size properties are checked on the way: the drillDown output rolls
back up to its input (sum over the days of a month == the month's cell, rel 1e-6)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    from olap_in_memory_b200 import Cube, GenericDimension, TimeDimension, _native, interop

    _native.init(0)
    lib = _native.lib()
    interop.use_torch_stream()
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dims = [GenericDimension(name, "root", [f"{name}{i}" for i in range(n)]) for name, n in (("a", 100), ("b", 100), ("c", 100))]
    dims += [TimeDimension("time", "month", "2010-01", "2010-10")]
    dims += [GenericDimension(name, "root", [f"{name}{i}" for i in range(10)]) for name in ("e", "f")]
    cube = Cube(dims)
    cube.createStoredMeasure("amount", {"time": "sum"}, "float32", 0)
    interop.values_tensor(cube.storedMeasures["amount"]).uniform_(1.0, 1000.0)
    interop.status_tensor(cube.storedMeasures["amount"]).fill_(2)
    torch.cuda.synchronize()
    rows = []

    def timed(label, fn, algo_bytes, cells_in, cells_out, keep=False):
        out = fn()
        ms = []
        for _ in range(args.reps):
            if not keep:
                del out
            out = fn()
            ms.append(lib.olap_last_op_ms())
        t = float(np.median(ms))
        row = {"op": label, "path": lib.olap_last_op_path().decode(), "cells_in": cells_in, "cells_out": cells_out, "kernel_ms": round(t, 3),
               "algorithmic_GB": round(algo_bytes / 1e9, 2), "GBs": round(algo_bytes / (t * 1e-3) / 1e9, 1),
               "frac_of_measured_hbm_peak": round(algo_bytes / (t * 1e-3) / 1e9 / peak, 3), "cells_out_per_s": cells_out / (t * 1e-3)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        return out

    n = cube.storeSize
    every_other = [f"a{i}" for i in range(0, 100, 2)]
    diced = timed("dice a -> every other item", lambda: cube.dice("a", "root", every_other), 5 * 2 * (n // 2), n, n // 2)
    rev = timed("reorderDimensions -> reversed", lambda: cube.reorderDimensions(["f", "e", "time", "c", "b", "a"]), 5 * 2 * n, n, n)
    ref = interop.values_tensor(cube.storedMeasures["amount"]).view(100, 100, 100, 10, 10, 10)
    got = interop.values_tensor(rev.storedMeasures["amount"]).view(10, 10, 10, 100, 100, 100)
    assert torch.equal(got[3, 7, 1, :, 42, 5], ref[5, 42, :, 1, 7, 3]) and torch.equal(got[:, 0, 9, 99, 0, 50], ref[50, 0, 99, 9, 0, :])
    del rev, got, ref
    n_d = diced.storeSize
    n_out = n_d // 10 * 304
    down = timed("drillDown time month -> day of the diced cube", lambda: diced.drillDown("time", "day"), 5 * (n_d + n_out), n_d, n_out)
    assert down.storeSize == n_out == 15_200_000_000
    # property at full size: days roll back up to their month (the reference's down-then-up round trip, test/cube-drilling.js:85-140)
    back = down.drillUp("time", "month")
    a = interop.values_tensor(back.storedMeasures["amount"])
    b = interop.values_tensor(diced.storedMeasures["amount"])
    rel = float(((a - b).abs() / b.abs()).max().item())
    st = interop.status_tensor(down.storedMeasures["amount"])
    flags = torch.unique(st[:: 1000003]).tolist()
    row = {"check": "drillUp(drillDown(x)) == x at 1.52e10 cells", "max_rel_err": rel, "status_flags_seen": flags,
           "roll_back_kernel_ms": round(lib.olap_last_op_ms(), 3)}
    assert rel <= 1e-6 and flags == [6], row  # SET | INTERPOLATED on every day cell
    rows.append(row)
    print(json.dumps(row), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
