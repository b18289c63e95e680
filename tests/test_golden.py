"""Committed golden vectors (tests/golden/store_vectors.json, made by
tests/golden/make_golden.py): both oracles on CPU, the CUDA store on the GPU."""
import numpy as np
import pytest

import cases
import golden_io
from oracle.c_oracle import COracleStore
from oracle.store_oracle import OracleStore

VECTORS = golden_io.load()
IDS = [f"{i}-{c['op']}" for i, (c, _) in enumerate(VECTORS)]


def _same_f64(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.all((np.isnan(a) & np.isnan(b)) | (a.view(np.uint64) == b.view(np.uint64))))


@pytest.mark.parametrize("case,expected", VECTORS, ids=IDS)
def test_python_oracle_reproduces_golden(case, expected):
    assert _same_f64(cases.run_case(case, OracleStore), expected), cases.describe(case)


@pytest.mark.parametrize("case,expected", VECTORS, ids=IDS)
def test_c_oracle_reproduces_golden(case, expected):
    assert _same_f64(cases.run_case(case, COracleStore), expected), cases.describe(case)


@pytest.mark.gpu
@pytest.mark.parametrize("case,expected", VECTORS, ids=IDS)
def test_gpu_store_reproduces_golden(case, expected):
    from olap_in_memory_b200.store import GpuStore

    got = cases.run_case(case, GpuStore)
    assert cases.bits_equal(got, expected.astype(np.float32)), cases.describe(case)
