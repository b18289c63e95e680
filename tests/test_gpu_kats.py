"""The reference's own test expectations (tests/kats.py), run on the device store
through the C ABI."""
import pytest

import kats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kat", kats.ALL_KATS, ids=lambda f: f.__name__)
def test_reference_kat_on_gpu(kat):
    from olap_in_memory_b200.store import GpuStore

    kat(GpuStore)
