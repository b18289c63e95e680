"""Pins the CPU oracle (oracle/store_oracle.py) on the reference's own test
expectations (tests/kats.py, transcribed from /root/reference/test/*.js)."""
import pytest

import kats
from oracle.store_oracle import OracleStore


@pytest.mark.parametrize("kat", kats.ALL_KATS, ids=lambda f: f.__name__)
def test_reference_kat_on_oracle(kat):
    kat(OracleStore)
