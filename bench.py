#!/usr/bin/env python
"""bench.py — drillUp throughput of the cube-transform hot path on B200.

A "step" is one pass of the hot path over one batch of synthetic input: the config-2
cube of BASELINE.json (time day x3652, 3 generic dims x32 -> 119 668 736 cells, 3 stored
measures with time rules sum / average / highest, 1 computed measure) is drilled up
day -> month (one batched olap_drill_up call for the 3 measures) and the computed
measure `(m_sum + m_avg) / m_max` is evaluated on the result by the JIT-fused kernel.

  value   : measure-cells/s (M x N_in / t) with the cube resident in HBM
  e2e     : the same metric through the public Cube API with HOST buffers: every step
            uploads the 3 measures from pinned host memory, runs the step and downloads
            the 4 result planes
  roofline: the drillUp kernel's algorithmic bytes / its CUDA-event duration, against
            the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the reference's algorithm (oracle/olap_oracle.c, a
            plain-C restatement of src/store/in-memory.js) on the host cores, on a
            bounded sample of the same workload.

Launch:  python bench.py [--gpus N --steps K --warmup W] ; for N > 1 under torchrun,
one rank per GPU, each rank holding its own config-2 cube (weak scaling: the path
shards on an outer axis with no data-path collective)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TIME_START, TIME_END = "2010-01-01", "2019-12-31"
GENERIC = 32
METHODS = ("sum", "average", "highest")
FORMULA = "(m_sum + m_avg) / m_max"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ------------------------------------------------------------------ synthetic data
def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synth(out, seed, measure, start=0):
    """SURVEY.md §8d generator: float32 values uniform in [1, 1000), never 0 / NaN."""
    n = out.size
    chunk = 1 << 24
    with np.errstate(over="ignore"):
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            idx = np.arange(start + lo, start + hi, dtype=np.uint64)
            h = splitmix64(idx ^ np.uint64(seed) ^ (np.uint64(measure) << np.uint64(40)))
            out[lo:hi] = (1.0 + (h >> np.uint64(40)).astype(np.float64) * (999.0 / (1 << 24))).astype(np.float32)
    return out


class Clocks(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML every ~2 ms while the
    timed regions run (the nvidia-smi CLI is too slow for sub-second regions)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.max_mhz = index, [], 0, None
        self.stop_flag = threading.Event()
        self.busy = threading.Event()  # set while a timed region is running

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:
            return
        while not self.stop_flag.is_set():
            if self.busy.is_set():
                try:
                    self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    self.reasons |= int(get_reasons(h))
                except Exception:
                    pass
            time.sleep(0.002)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": self.max_mhz, "reasons": [k for k, v in bits.items() if self.reasons & v],
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle port)
def cpu_port_run(steps, warmup, threads, inner):
    """The reference's drillUp algorithm on host cores: each thread owns an independent
    sub-cube [3652 days x inner] per measure and rolls it up day -> month.  inner = 4096
    gives 14.96 M cells per measure, just under the 2^24-entry limit of the JS Map the
    reference keeps its cells in."""
    from olap_in_memory_b200 import TimeDimension
    from oracle.c_oracle import COracleStore

    day = TimeDimension("time", "day", TIME_START, TIME_END)
    month_map = np.asarray(day.getGroupIndexFromRootIndexMap("month"), dtype=np.int32)
    C, P, I = day.numItems, 120, inner
    n = C * I
    ident = np.arange(I, dtype=np.int32)
    stores = []
    for t in range(threads):
        per = []
        for m, _ in enumerate(METHODS):
            s = COracleStore(n, "float32", 0.0)
            s.set_data_f32(synth(np.empty(n, np.float32), 1 + t, m))
            per.append(s)
        stores.append(per)

    def work(t):
        for s, method in zip(stores[t], METHODS):
            s.drillUp_lowered([C, I], [P, I], [month_map, ident], method)

    def one_step():
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    cells = threads * len(METHODS) * n
    sample = (f"drillUp day->month of [3652 x {I}] cells x {len(METHODS)} measures (sum/average/highest) per thread, "
              f"{threads} independent sub-cubes, C port of in-memory.js:265-334 (ordered hash map)")
    return cells / dt, dt * 1e3, sample


def run_reference(args, rank):
    if rank != 0:
        return
    threads = max(1, min(os.cpu_count() or 1, 32))
    value, ms, sample = cpu_port_run(max(1, args.steps), max(0, min(args.warmup, 1)), threads, 512)
    line = {
        "impl": "reference", "metric": "drillUp input measure-cells/s", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config():
    return {"workload": "config 2: cube [time day x3652, g1 x32, g2 x32, g3 x32] = 119668736 cells, 3 stored float32 "
                        "measures (time rules sum/average/highest) + 1 computed measure; drillUp time day->month "
                        "-> 3932160 cells; per-measure status plane (5 B/cell)",
            "cells_in": 3652 * GENERIC ** 3, "cells_out": 120 * GENERIC ** 3, "measures": 3, "computed": FORMULA,
            "l2": "inputs (2.4 GB) exceed the 126 MB L2, no flush needed", "sharding": "one config-2 cube per GPU"}


# ------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from olap_in_memory_b200 import Cube, GenericDimension, TimeDimension
    from olap_in_memory_b200 import _native as N

    torch.cuda.set_device(local_rank)
    N.init(local_rank)
    lib = N.lib()
    stream = torch.cuda.Stream()
    N.check(lib.olap_set_stream(stream.cuda_stream))

    dims = [TimeDimension("time", "day", TIME_START, TIME_END)] + [
        GenericDimension(f"g{k}", "root", [f"g{k}-{i}" for i in range(GENERIC)]) for k in (1, 2, 3)]
    cube = Cube(dims)
    names = ["m_sum", "m_avg", "m_max"]
    for name, rule in zip(names, METHODS):
        cube.createStoredMeasure(name, {"time": rule}, "float32", 0)
    cube.createComputedMeasure("ratio", FORMULA)
    n_in = cube.storeSize
    n_out = 120 * GENERIC ** 3

    # pinned host buffers (the e2e path copies from / to these every step)
    import ctypes as C

    def pinned(n_floats):
        p = C.c_void_p()
        N.check(lib.olap_host_alloc(n_floats * 4, C.byref(p)))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n_floats,)), p

    host_in, host_out, keep = [], [], []
    for m, name in enumerate(names):
        arr, p = pinned(n_in)
        synth(arr, 1 + rank, m)
        host_in.append(arr)
        keep.append(p)
        cube.setData(name, arr)
    for _ in range(4):
        arr, p = pinned(n_out)
        host_out.append(arr)
        keep.append(p)

    def step_device():
        rolled = cube.drillUp("time", "month")
        ms_up = None
        if step_device.measure:
            ms_up = lib.olap_last_op_ms()
            step_device.path = lib.olap_last_op_path().decode()
        ratio = rolled.evaluateToStore("ratio")
        return rolled, ratio, ms_up

    step_device.measure = False

    def step_e2e():
        for name, arr in zip(names, host_in):
            cube.setData(name, arr)
        rolled, ratio, _ = step_device()
        for name, out in zip(names, host_out):
            N.check(lib.olap_store_download_f32(rolled.storedMeasures[name]._h, out.ctypes.data, n_out))
        N.check(lib.olap_store_download_f32(ratio._h, host_out[3].ctypes.data, n_out))
        return float(host_out[3][0])

    clocks = Clocks(local_rank)
    clocks.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        timed.launches = lib.olap_kernel_launches()
        clocks.busy.set()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        clocks.busy.clear()
        timed.launches = lib.olap_kernel_launches() - timed.launches
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # correctness guard before timing: month sums of one column against numpy (float64)
    rolled, ratio, _ = step_device()
    col = host_in[0].reshape(3652, -1)[:, 12345].astype(np.float64)
    month_map = np.asarray(dims[0].getGroupIndexFromRootIndexMap("month"))
    want = np.zeros(120)
    np.add.at(want, month_map, col)
    got = rolled.storedMeasures["m_sum"].data_f32().reshape(120, -1)[:, 12345]
    assert np.array_equal(got, want.astype(np.float32)), "bench: drillUp result does not match the float64 reference"
    del rolled, ratio

    N.check(lib.olap_set_async(1))
    ms_value = timed(step_device, args.steps, args.warmup)
    launches = timed.launches  # kernels of libolapgpu.so launched inside the timed region
    N.check(lib.olap_set_async(0))

    # per-launch duration of the dominant kernel, CUDA events on the launching stream
    step_device.measure = True
    ups = []
    for _ in range(max(3, args.steps)):
        _, _, ms_up = step_device()
        ups.append(ms_up)
    step_device.measure = False
    up_ms = float(np.mean(ups))
    path = step_device.path

    ms_e2e = timed(step_e2e, max(1, min(args.steps, 5)), 1)
    clk = clocks.summary()

    measures = len(names)
    cells_per_step = measures * n_in
    value = world * cells_per_step / (ms_value * 1e-3)
    e2e_value = world * cells_per_step / (ms_e2e * 1e-3)
    bytes_per_cell = 5  # float32 cell + status byte, per measure
    algo_bytes = bytes_per_cell * measures * (n_in + n_out)
    achieved = algo_bytes / (up_ms * 1e-3) / 1e9
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured"
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "drillup_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = 1
            v, _, sample = cpu_port_run(1, 0, threads, GENERIC * GENERIC * 4)
            cpu = {"value": v, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample}
        line = {
            "metric": "drillUp input measure-cells/s", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(),
            "e2e": {"value": e2e_value, "unit": "cells/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": world * measures * n_in * 4, "d2h_bytes_per_step": world * 4 * n_out * 4},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": path, "kernel_ms": up_ms, "algorithmic_bytes": algo_bytes,
                         "peak_source": peak_src},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
