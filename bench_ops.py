#!/usr/bin/env python
"""Per-kernel sweep (developer tool, not the driver contract): every transform of the
path on the shape classes of SURVEY.md Appendix B, timed with CUDA events on the launching
stream, reported as algorithmic GB/s against the measured HBM peak.
  python bench_ops.py [--scale 1.0] [--only drillup,dice,...] [--status 0|1]"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="")
    ap.add_argument("--status", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch

    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200 import interop
    from olap_in_memory_b200.store import GpuStore
    from olap_in_memory_b200 import TimeDimension

    N.init(0)
    lib = N.lib()
    interop.use_torch_stream()
    GpuStore.WITH_STATUS = bool(args.status)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    S = 1 if args.status else 0
    only = set(filter(None, args.only.split(",")))
    results = []

    def store(n, default=0.0, fill=1.0, typ="float32"):
        s = GpuStore(n, typ, default)
        v = interop.values_tensor(s)
        v.uniform_(1.0, 1000.0)
        if fill < 1.0:
            mask = torch.rand(n, device="cuda") >= fill
            v[mask] = float("nan") if default != default else 0.0
            del mask
        st = interop.status_tensor(s)
        if st is not None:
            st.fill_(2)
            if fill < 1.0:
                pres = (v == v) if default != default else (v != 0)
                st[~pres] = 1
        torch.cuda.synchronize()
        return s

    def run(name, fn, algo_bytes, cells):
        if only and not any(name.startswith(o) for o in only):
            return
        out = fn()
        del out
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.reps):
            out = fn()
            ms.append(lib.olap_last_op_ms())
            path = lib.olap_last_op_path().decode()
            del out
        t = float(np.median(ms))
        gbs = algo_bytes / (t * 1e-3) / 1e9
        row = {"op": name, "path": path, "ms": round(t, 4), "GBs": round(gbs, 1), "frac": round(gbs / peak, 3),
               "cells_per_s": cells / (t * 1e-3)}
        results.append(row)
        print(json.dumps(row), flush=True)

    day = TimeDimension("time", "day", "2010-01-01", "2019-12-31")
    month = np.asarray(day.getGroupIndexFromRootIndexMap("month"), np.int32)
    year = np.asarray(day.getGroupIndexFromRootIndexMap("year"), np.int32)
    ident = lambda n: np.arange(n, dtype=np.int32)
    sc = args.scale
    B = 4 + S

    # ---- drillUp ---------------------------------------------------------------
    I = int(32768 * sc)
    for method in ("sum", "average", "highest", "first"):
        for default in (0.0, float("nan")):
            tag = "nan" if default != default else "zero"
            s = store(3652 * I, default)
            run(f"drillup/time-outer day->month {method} {tag} [1,3652,{I}]",
                lambda: GpuStore.drillUp_lowered([s], [3652, I], [120, I], [month, ident(I)], [method]),
                B * (3652 + 120) * I, 3652 * I)
            del s
    # the same rollup of a store whose status plane follows from its values (what setData leaves behind): the
    # plane is not read, 4 bytes per input cell + 5 per output cell
    for method in ("sum", "average", "highest"):
        s = store(3652 * I, 0.0)
        s.canonicalise()
        run(f"drillup/derived-status day->month {method} zero [1,3652,{I}]",
            lambda: GpuStore.drillUp_lowered([s], [3652, I], [120, I], [month, ident(I)], [method]),
            (4 * 3652 + B * 120) * I, 3652 * I)
        del s
    s = store(3652 * I, 0.0)
    s.canonicalise()
    run(f"drillup/derived-status time-inner day->month sum [{I},3652,1]",
        lambda: GpuStore.drillUp_lowered([s], [I, 3652], [I, 120], [ident(I), month], ["sum"]), (4 * 3652 + B * 120) * I, 3652 * I)
    del s
    s = store(3652 * I, 0.0, 0.25)
    run(f"drillup/time-outer day->month sum zero fill=0.25 [1,3652,{I}]",
        lambda: GpuStore.drillUp_lowered([s], [3652, I], [120, I], [month, ident(I)], ["sum"]), B * (3652 + 120) * I, 3652 * I)
    run(f"drillup/time-outer day->year sum [1,3652,{I}]",
        lambda: GpuStore.drillUp_lowered([s], [3652, I], [10, I], [year, ident(I)], ["sum"]), B * (3652 + 10) * I, 3652 * I)
    run(f"drillup/time-outer day->all sum [1,3652,{I}]",
        lambda: GpuStore.drillUp_lowered([s], [3652, I], [1, I], [np.zeros(3652, np.int32), ident(I)], ["sum"]),
        B * (3652 + 1) * I, 3652 * I)
    run(f"drillup/time-inner day->month sum [{I},3652,1]",
        lambda: GpuStore.drillUp_lowered([s], [I, 3652], [I, 120], [ident(I), month], ["sum"]), B * (3652 + 120) * I, 3652 * I)
    run(f"drillup/time-inner day->all sum [{I},3652,1]",
        lambda: GpuStore.drillUp_lowered([s], [I, 3652], [I, 1], [ident(I), np.zeros(3652, np.int32)], ["sum"]),
        B * (3652 + 1) * I, 3652 * I)
    # long rows, few parents (drillup/long): collapse of the whole cube, 100 000 customers -> 8 segments
    nL = 3652 * I
    run(f"drillup/long collapse sum [1,{nL},1]",
        lambda: GpuStore.drillUp_lowered([s], [nL], [1], [np.zeros(nL, np.int32)], ["sum"]), B * (nL + 1), nL)
    run(f"drillup/long collapse first [1,{nL},1]",
        lambda: GpuStore.drillUp_lowered([s], [nL], [1], [np.zeros(nL, np.int32)], ["first"]), B * (nL + 1), nL)
    del s
    Oc, Cc = nL // 100000, 100000
    s = store(Oc * Cc, 0.0)
    seg8 = np.random.default_rng(0).integers(0, 8, Cc).astype(np.int32)
    run(f"drillup/long customers->segment sum [{Oc},{Cc},1]",
        lambda: GpuStore.drillUp_lowered([s], [Oc, Cc], [Oc, 8], [ident(Oc), seg8], ["sum"]), B * (Cc + 8) * Oc, Cc * Oc)
    run(f"drillup/long customers->segment average [{Oc},{Cc},1]",
        lambda: GpuStore.drillUp_lowered([s], [Oc, Cc], [Oc, 8], [ident(Oc), seg8], ["average"]), B * (Cc + 8) * Oc, Cc * Oc)
    s.canonicalise()  # what setData leaves behind: the status plane follows from the values and is not read
    run(f"drillup/derived-status customers->segment sum [{Oc},{Cc},1]",
        lambda: GpuStore.drillUp_lowered([s], [Oc, Cc], [Oc, 8], [ident(Oc), seg8], ["sum"]), (4 * Cc + B * 8) * Oc, Cc * Oc)
    run(f"drillup/derived-status customers->segment average [{Oc},{Cc},1]",
        lambda: GpuStore.drillUp_lowered([s], [Oc, Cc], [Oc, 8], [ident(Oc), seg8], ["average"]), (4 * Cc + B * 8) * Oc, Cc * Oc)
    seg8m = np.sort(seg8)
    run(f"drillup/derived-status customers->segment (sorted map) first [{Oc},{Cc},1]",
        lambda: GpuStore.drillUp_lowered([s], [Oc, Cc], [Oc, 8], [ident(Oc), seg8m], ["first"]), (4 * Cc + B * 8) * Oc, Cc * Oc)
    del s
    # univac-style: 10-item generic dims, 1e9 cells at scale 1 (identity axes are split so
    # that no map is longer than 1e4 entries)
    h = int(round(1e4 * sc ** 0.5))
    s = store(h * h * 10, 0.0)
    zeros10 = np.zeros(10, np.int32)
    parity = (np.arange(10) % 2).astype(np.int32)
    nU = h * h * 10
    run(f"drillup/inner dim9->all sum [{h * h},10,1]",
        lambda: GpuStore.drillUp_lowered([s], [h, h, 10], [h, h, 1], [ident(h), ident(h), zeros10], ["sum"]),
        B * 11 * h * h, nU)
    run(f"drillup/inner dim9->parity sum [{h * h},10,1]",
        lambda: GpuStore.drillUp_lowered([s], [h, h, 10], [h, h, 2], [ident(h), ident(h), parity], ["sum"]),
        B * 12 * h * h, nU)
    run(f"drillup/mid dim5->all sum [{h},10,{h}]",
        lambda: GpuStore.drillUp_lowered([s], [h, 10, h], [h, 1, h], [ident(h), zeros10, ident(h)], ["sum"]),
        B * 11 * h * h, nU)
    run(f"drillup/mid dim5->parity sum [{h},10,{h}]",
        lambda: GpuStore.drillUp_lowered([s], [h, 10, h], [h, 2, h], [ident(h), parity, ident(h)], ["sum"]),
        B * 12 * h * h, nU)
    run(f"drillup/outer dim0->all sum [1,10,{h * h}]",
        lambda: GpuStore.drillUp_lowered([s], [10, h, h], [1, h, h], [zeros10, ident(h), ident(h)], ["sum"]),
        B * 11 * h * h, nU)
    del s

    # ---- dice / reorder / drillDown on a config-3 style cube -----------------------------
    a = int(100 * sc) if sc < 1 else 100
    dims = [a, 100, 100, 10, 10, 10]
    n = int(np.prod(dims))
    s = store(n, 0.0)
    keep = [np.arange(0, a, 2, dtype=np.int32)] + [ident(d) for d in dims[1:]]
    run(f"dice/outer every-other a {dims}", lambda: GpuStore.dice_lowered([s], dims, keep), B * 2 * (n // 2), n // 2)
    keep_c = [ident(d) for d in dims]
    keep_c[2] = np.arange(0, 100, 2, dtype=np.int32)
    run(f"dice/mid every-other c {dims}", lambda: GpuStore.dice_lowered([s], dims, keep_c), B * 2 * (n // 2), n // 2)
    keep_f = [ident(d) for d in dims]
    keep_f[5] = np.arange(0, 10, 2, dtype=np.int32)
    run(f"dice/inner every-other f {dims}", lambda: GpuStore.dice_lowered([s], dims, keep_f), B * 2 * (n // 2), n // 2)
    keep_r = [ident(d) for d in dims]
    keep_r[3] = np.arange(2, 8, dtype=np.int32)
    run(f"dice/range t[2:8] {dims}", lambda: GpuStore.dice_lowered([s], dims, keep_r), B * 2 * (n * 6 // 10), n * 6 // 10)
    run(f"reorder/reverse axes {dims}", lambda: GpuStore.reorder_lowered([s], dims, [5, 4, 3, 2, 1, 0]), B * 2 * n, n)
    if S:
        s.canonicalise()  # status plane derived from the values: gathers and the pair transpose write it without reading it
        run(f"dice/derived-status outer every-other a {dims}", lambda: GpuStore.dice_lowered([s], dims, keep), (4 + B) * (n // 2), n // 2)
        run(f"reorder/derived-status reverse axes {dims}", lambda: GpuStore.reorder_lowered([s], dims, [5, 4, 3, 2, 1, 0]), (4 + B) * n, n)
        run(f"reorder/derived-status swap inner two {dims}", lambda: GpuStore.reorder_lowered([s], dims, [0, 1, 2, 3, 5, 4]), (4 + B) * n, n)
        run(f"reorder/derived-status rotate inner to front {dims}", lambda: GpuStore.reorder_lowered([s], dims, [5, 0, 1, 2, 3, 4]), (4 + B) * n, n)
        interop.status_tensor(s)  # a mutable pointer was handed out: back to the loaded plane for the rows below
    run(f"reorder/swap outer two {dims}", lambda: GpuStore.reorder_lowered([s], dims, [1, 0, 2, 3, 4, 5]), B * 2 * n, n)
    run(f"reorder/swap inner two {dims}", lambda: GpuStore.reorder_lowered([s], dims, [0, 1, 2, 3, 5, 4]), B * 2 * n, n)
    run(f"reorder/rotate inner to front {dims}", lambda: GpuStore.reorder_lowered([s], dims, [5, 0, 1, 2, 3, 4]), B * 2 * n, n)
    rdims = dims[::-1]  # the mirror image: a short OUTER axis becomes the innermost one
    run(f"reorder/rotate front to inner {rdims}", lambda: GpuStore.reorder_lowered([s], rdims, [1, 2, 3, 4, 5, 0]), B * 2 * n, n)
    del s
    # drillDown month -> day (10 months -> 304 days) with I = 100 and I = 1
    mdim = TimeDimension("time", "month", "2010-01", "2010-10")
    ddim = mdim.drillDown("day")
    m2d = np.asarray(ddim.getGroupIndexFromRootIndexMap("month"), np.int32)
    Od = int(5e3 * sc)
    s = store(Od * 10 * 100, 0.0)
    run(f"drilldown/mid month->day float [{Od},10->304,100]",
        lambda: GpuStore.drillDown_lowered([s], [Od, 10, 100], [Od, 304, 100], [ident(Od), m2d, ident(100)], ["sum"]),
        B * (10 + 304) * Od * 100, 304 * Od * 100)
    run(f"drilldown/inner month->day float [{Od * 100},10->304,1]",
        lambda: GpuStore.drillDown_lowered([s], [Od * 100, 10], [Od * 100, 304], [ident(Od * 100), m2d], ["sum"]),
        B * (10 + 304) * Od * 100, 304 * Od * 100)
    del s
    Ob = int(5e4 * sc)  # the config-3 sized interpolation: 5e7 cells -> 1.52e9 cells
    s = store(Ob * 10 * 100, 0.0)
    run(f"drilldown/mid month->day float [{Ob},10->304,100]",
        lambda: GpuStore.drillDown_lowered([s], [Ob, 10, 100], [Ob, 304, 100], [ident(Ob), m2d, ident(100)], ["sum"]),
        B * (10 + 304) * Ob * 100, 304 * Ob * 100)
    del s
    si = store(Od * 10 * 100, 0.0, typ="uint32")
    run(f"drilldown/mid month->day uint32 [{Od},10->304,100]",
        lambda: GpuStore.drillDown_lowered([si], [Od, 10, 100], [Od, 304, 100], [ident(Od), m2d, ident(100)], ["sum"]),
        B * (10 + 304) * Od * 100, 304 * Od * 100)
    del si

    # ---- computed measure, total ----------------------------------------------------------
    from olap_in_memory_b200.parser import getParser

    nI = int(119668736 * sc)
    ins = [store(nI, 0.0) for _ in range(3)]
    expr = getParser().parse("(a + b) / c")
    run(f"eval/(a+b)/c -> store [{nI}]",
        lambda: GpuStore.evaluate_to_store(expr, ["a", "b", "c"], ins, {}), 4 * 3 * nI + (4 + S) * nI, nI)
    expr2 = getParser().parse("a || b * 2 - isNaN(c)")
    run(f"eval/a||b*2-isNaN(c) -> store [{nI}]",
        lambda: GpuStore.evaluate_to_store(expr2, ["a", "b", "c"], ins, {}), 4 * 3 * nI + (4 + S) * nI, nI)
    # ---- cube-to-cube load (N2) and sparse export / import (N3) ----------------------------
    dl = [int(1000 * sc) if sc < 1 else 1000, 100, 100, 10]
    nl = int(np.prod(dl))
    his = store(nl, 0.0, fill=0.5)
    mine = GpuStore(nl * 2, "float32", 0.0)  # twice the items on the outer axis
    to_mine = [np.arange(0, 2 * dl[0], 2, dtype=np.int32)] + [ident(d) for d in dl[1:]]

    def timed_host(name, fn, algo_bytes, cells, kernel_bracket=False):
        if only and not any(name.startswith(o) for o in only):
            return
        import time

        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        t = float(np.median(ts))
        row = {"op": name, "path": lib.olap_last_op_path().decode(), "ms": round(t, 4), "GBs": round(algo_bytes / (t * 1e-3) / 1e9, 1),
               "frac": round(algo_bytes / (t * 1e-3) / 1e9 / peak, 3), "cells_per_s": cells / (t * 1e-3), "timing": "host wall-clock, synchronous call"}
        if kernel_bracket:  # the op's kernels alone (event bracket inside the library)
            k = float(lib.olap_last_op_ms())
            row["kernel_ms"] = round(k, 4)
            row["kernel_frac"] = round(algo_bytes / (k * 1e-3) / 1e9 / peak, 3) if k > 0 else None
        results.append(row)
        print(json.dumps(row), flush=True)

    timed_host(f"load/scatter every-other outer item {dl} -> x2", lambda: mine.load_lowered(his, [2 * dl[0]] + dl[1:], dl, to_mine),
               B * 2 * nl, nl, kernel_bracket=True)
    timed_host(f"sparse/export fill 0.5 [{nl}] (12 B per set cell to the host)", lambda: his.export_sparse(), B * nl + 12 * (nl // 2), nl)
    del his, mine
    if not only or "total" in only:
        import time

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ins[0].total
        t = (time.perf_counter() - t0) / 5 * 1e3
        k = float(lib.olap_last_op_ms())
        row = {"op": f"total [{nI}] (host-timed, includes sync + D2H of 16 B)", "ms": round(t, 4),
               "GBs": round(4 * nI / t / 1e6, 1), "frac": round(4 * nI / t / 1e6 / peak, 3),
               "kernel_ms": round(k, 4), "kernel_frac": round(4 * nI / k / 1e6 / peak, 3) if k > 0 else None}
        results.append(row)
        print(json.dumps(row), flush=True)
    if not only or "boundary" in only:
        # the data boundary with PAGEABLE host arrays (what a typed array of the addon is): csrc/host_pipe.cuh
        import time
        import ctypes as C

        nB = nI
        host = np.random.default_rng(1).random(nB, dtype=np.float32) + 1.0
        sB = GpuStore(nB, "float32", 0.0)
        out_buf = np.zeros(nB, dtype=np.float32)  # touched once: no first-touch page faults in the timing

        def wall(fn, reps=3):
            fn()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return float(np.median(ts))

        t_up = wall(lambda: sB.set_data_f32(host))
        t_down = wall(lambda: N.check(lib.olap_store_download_f32(sB._h, out_buf.ctypes.data, nB)))
        t_fresh = wall(lambda: sB.data_f32())
        assert np.array_equal(out_buf, host)
        for name, t in (("boundary/upload pageable Float32Array", t_up), ("boundary/download into a touched pageable array", t_down),
                        ("boundary/download into a fresh pageable array (first touch included)", t_fresh)):
            row = {"op": f"{name} [{nB}]", "ms": round(t * 1e3, 3), "GBs": round(4 * nB / t / 1e9, 2),
                   "copy_threads": os.environ.get("OLAP_COPY_THREADS", "default"), "host_pipe": os.environ.get("OLAP_HOST_PIPE", "1")}
            results.append(row)
            print(json.dumps(row), flush=True)
        del sB
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
