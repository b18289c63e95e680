"""Sharded cube on real GPUs over NCCL (needs >= 2 devices; skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_cube_over_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "gpu_sharded_check.py")]
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "failures=0" in proc.stdout


@pytest.mark.parametrize("prefix", [1, 2])
def test_single_rank_sharded_cube_on_the_device(prefix):
    """World of one (no process group): every ShardedCube transform — including dice and
    drillDown of the sharded dimensions, which lower to row-axis gathers — on the device
    store against ONE oracle cube."""
    import math

    import numpy as np

    from olap_in_memory_b200 import Cube
    from olap_in_memory_b200.sharded import ShardedCube
    from oracle.store_oracle import OracleStore
    from test_sharded_gloo import (_collect, _dims, _fill, _time_first_collect, _time_first_dims, _time_first_fill)

    for default in (0.0, math.nan):
        cube, ref = ShardedCube(_dims(), prefix=prefix), Cube(_dims(), OracleStore)
        _fill(cube, default)
        _fill(ref, default)
        got, want = _collect(cube, list(cube.storedMeasures)), _collect(ref, ref.storedMeasureIds)
        assert got.keys() == want.keys()
        for key in want:
            if not isinstance(key, str) and key[0] == "chain" and key[2] in ("m_first", "m_last"):
                continue  # declared divergence of the reference's Map order
            assert np.allclose(got[key], want[key], rtol=1e-6, atol=0, equal_nan=True), (prefix, default, key)
    cube, ref = ShardedCube(_time_first_dims(), prefix=prefix), Cube(_time_first_dims(), OracleStore)
    got = _time_first_collect(cube, _time_first_fill(cube, 0.0))
    want = _time_first_collect(ref, _time_first_fill(ref, 0.0))
    for key in want:
        assert np.allclose(got[key], want[key], rtol=1e-6, atol=0, equal_nan=True), (prefix, key)
