"""The C oracle (oracle/olap_oracle.c) must agree bit for bit — doubles, and Map
insertion order — with the Python oracle, which is pinned on the reference's own
tests (tests/test_oracle_kats.py).  CPU only."""
import itertools

import numpy as np
import pytest

import cases
from oracle.c_oracle import COracleStore
from oracle.store_oracle import OracleStore

ALL = list(itertools.chain(cases.drillup_cases(), cases.drilldown_cases(), cases.dice_cases(),
                           cases.reorder_cases(), cases.load_cases(), cases.load_linear_cases()))


def _same_f64(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.all((np.isnan(a) & np.isnan(b)) | (a.view(np.uint64) == b.view(np.uint64))))


@pytest.mark.parametrize("case", ALL, ids=[f"{i}-{c['op']}" for i, c in enumerate(ALL)])
def test_c_oracle_matches_python_oracle(case):
    assert _same_f64(cases.run_case(case, COracleStore), cases.run_case(case, OracleStore)), cases.describe(case)


def test_insertion_order_and_delete_reinsert():
    py, c = OracleStore(8, "float32", 0), COracleStore(8, "float32", 0)
    for store in (py, c):
        for idx, v in ((5, 1.0), (2, 2.0), (7, 3.0), (2, 9.0), (5, 0.0), (5, 4.0)):
            store.setValue(idx, v)
    keys, vals = c.entries()
    assert keys.tolist() == list(py._dataMap.keys()) == [2, 7, 5]
    assert vals.tolist() == list(py._dataMap.values()) == [9.0, 3.0, 4.0]
    assert c.total == py.total == 16.0


def test_first_last_follow_map_order_not_index_order():
    """dice(reorder=true) keeps the OLD insertion order."""
    for cls in (OracleStore, COracleStore):
        s = cls(3, "float32", 0)
        s.data = [10.0, 20.0, 30.0]
        diced = s.dice_lowered([3], [[2, 0, 1]])
        assert np.asarray(diced.data).tolist() == [30.0, 10.0, 20.0]
        first = diced.drillUp_lowered([3], [1], [np.zeros(3, np.int32)], "first")
        last = diced.drillUp_lowered([3], [1], [np.zeros(3, np.int32)], "last")
        assert np.asarray(first.data).tolist() == [10.0] and np.asarray(last.data).tolist() == [30.0]


def test_uint16_contribution_wrap():
    """65 536 set children wrap the Uint16 count to 0."""
    s = COracleStore(65536, "float32", 0)
    s.set_data_f32(np.full(65536, 2.0, np.float32))
    out = s.drillUp_lowered([65536], [1], [np.zeros(65536, np.int32)], "average")
    assert out.data == [131072.0]
