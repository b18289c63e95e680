#!/usr/bin/env python
"""Pure-read, pure-write and copy bandwidth of the HBM on this GPU (torch elementwise kernels, CUDA
events, best of 10): the ceilings a read-dominated (drillUp), a write-dominated (drillDown) and a
balanced (dice, reorder) kernel can reach, beside the copy figure in MEASURED_PEAKS.json."""
import json

import torch

n = 1 << 30  # 4 GiB of float32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")


def best(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


x.normal_()
out = {
    "write_GBs (fill_ of 4 GiB)": 4 * n / best(lambda: x.fill_(1.5)) / 1e6,
    "write_GBs (cudaMemset 4 GiB)": 4 * n / best(lambda: x.zero_()) / 1e6,
    "read_GBs (sum of 4 GiB)": 4 * n / best(lambda: x.sum()) / 1e6,
    "copy_GBs (read + write, 4 GiB -> 4 GiB)": 8 * n / best(lambda: y.copy_(x)) / 1e6,
    "write_1_read_30 (y[:n//32] = x[:n//32]; plus fill of the rest is not fused: skipped)": None,
}
print(json.dumps(out, indent=1))
