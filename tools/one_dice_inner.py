"""One dice of the innermost axis of the config-3 cube (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from olap_in_memory_b200 import _native as N, interop
from olap_in_memory_b200.store import GpuStore

N.init(0); interop.use_torch_stream()
lens = [100, 100, 100, 10, 10, 10]
n = int(np.prod(lens))
s = GpuStore(n, "float32", 0.0)
interop.values_tensor(s).uniform_(1.0, 1000.0)
interop.status_tensor(s).fill_(2)
keeps = [np.arange(d, dtype=np.int32) for d in lens[:-1]] + [np.arange(0, 10, 2, dtype=np.int32)]
for _ in range(3):
    out = GpuStore.dice_lowered([s], lens, keeps)
    print(N.lib().olap_last_op_path().decode(), N.lib().olap_last_op_ms())
    del out
