"""Out-of-bounds writes: compute-sanitizer is closed on the GPU pool, so the library's own
guard regions (OLAP_GUARD=1: 256 poisoned bytes after every plane, checked at destroy) are
exercised over the golden vectors and the seeded parity cases in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import gc, itertools, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import cases, golden_io
from olap_in_memory_b200 import _native
from olap_in_memory_b200.store import GpuStore
n = 0
for with_status in (True, False):
    GpuStore.WITH_STATUS = with_status
    for case in itertools.chain((c for c, _ in golden_io.load()), cases.drillup_cases(), cases.drillup_long_cases(), cases.drilldown_cases(),
                                cases.dice_cases(), cases.reorder_cases(), cases.load_cases(), cases.load_linear_cases()):
        cases.run_case(case, GpuStore)
        n += 1
gc.collect()
print("cases", n, "guard_violations", _native.lib().olap_guard_violations())
"""


def test_no_write_lands_in_a_guard_region():
    env = dict(os.environ, OLAP_GUARD="1")
    proc = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, os.path.join(ROOT, "tests"))], env=env,
                          capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "guard_violations 0" in proc.stdout, proc.stdout[-500:]
