"""One drillUp of the customers -> segment shape (for ncu captures): python tools/one_lanes.py [sum|average|first] [derived|loaded]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from olap_in_memory_b200 import _native as N, interop
from olap_in_memory_b200.store import GpuStore

method = sys.argv[1] if len(sys.argv) > 1 else "sum"
derived = (sys.argv[2] if len(sys.argv) > 2 else "derived") == "derived"
N.init(0); interop.use_torch_stream()
Oc, Cc = 1196, 100000
s = GpuStore(Oc * Cc, "float32", 0.0)
interop.values_tensor(s).uniform_(1.0, 1000.0)
if derived:
    s.canonicalise()
else:
    interop.status_tensor(s).fill_(2)
seg8 = np.random.default_rng(0).integers(0, 8, Cc).astype(np.int32)
for _ in range(4):
    out = GpuStore.drillUp_lowered([s], [Oc, Cc], [Oc, 8], [np.arange(Oc, dtype=np.int32), seg8], [method])
    print(N.lib().olap_last_op_path().decode(), N.lib().olap_last_op_ms())
    del out
