"""The C oracle behind the Cube API must pass the reference's own test expectations too
(it is the CPU column of bench_cube_benchmark.py)."""
import pytest

import kats
from oracle.c_oracle import COracleStore


@pytest.mark.parametrize("kat", kats.ALL_KATS, ids=lambda f: f.__name__)
def test_reference_kat_on_c_oracle(kat):
    kat(COracleStore)
