"""Cube-level fusion of multi-dimension rollups (cube.py `_remove_fused`): with a store
class that opts in, collapse / aggregateByDimensions / keepDimensions / removeDimensions
must return exactly what the reference's chain of removeDimension calls returns
(cube.js:320-324, 560-566, 890-908), and must fall back to the chain whenever a rule is
order-dependent."""
import math

import numpy as np
import pytest

from olap_in_memory_b200 import Cube, GenericDimension, TimeDimension
from oracle.store_oracle import OracleStore


class FusedOracleStore(OracleStore):
    FUSED_ROLLUPS = True
    calls = 0

    @staticmethod
    def drillUp_lowered(stores, old_len, new_len, maps, methods):
        FusedOracleStore.calls += 1
        return OracleStore.drillUp_lowered(stores, old_len, new_len, maps, methods)


def _cube(cls, measures):
    cube = Cube([TimeDimension("time", "month", "2010-01", "2010-04"),
                 GenericDimension("location", "city", ["paris", "toledo", "tokyo"]),
                 GenericDimension("kind", "k", ["a", "b"]),
                 GenericDimension("colour", "c", ["r", "g", "b", "y", "k"])], cls)
    rng = np.random.default_rng(9)
    for name, rules, default in measures:
        cube.createStoredMeasure(name, rules, "float32", default)
        data = rng.integers(-5, 20, cube.storeSize).astype(float)
        data[rng.random(cube.storeSize) < 0.4] = default
        cube.setData(name, data.tolist())
    cube.createComputedMeasure("both", " + ".join(m[0] for m in measures))
    return cube


FUSABLE = [("m_sum", {}, 0), ("m_hi", {d: "highest" for d in ("time", "location", "kind", "colour")}, math.nan),
           ("m_lo", {d: "lowest" for d in ("time", "location", "kind", "colour")}, 0)]
OPS = [lambda c: c.collapse(), lambda c: c.aggregateByDimensions(["kind"]), lambda c: c.keepDimensions(["time", "colour"]),
       lambda c: c.removeDimensions(["colour", "time"]), lambda c: c.removeDimensions(["location", "kind"]),
       lambda c: c.keepDimensions(["location"]), lambda c: c.project(["colour", "time"])]


def _same(a, b):
    assert a.dimensionIds == b.dimensionIds
    assert a.storedMeasuresRules == b.storedMeasuresRules
    for m in a.storedMeasureIds + a.computedMeasureIds:
        x, y = np.asarray(a.getData(m), float), np.asarray(b.getData(m), float)
        assert np.array_equal(x, y, equal_nan=True), m


@pytest.mark.parametrize("op", range(len(OPS)))
def test_fused_equals_chain(op):
    fused, plain = _cube(FusedOracleStore, FUSABLE), _cube(OracleStore, FUSABLE)
    FusedOracleStore.calls = 0
    _same(OPS[op](fused), OPS[op](plain))
    assert FusedOracleStore.calls >= 1  # the fused path ran (one lowered call per run of adjacent dimensions)


@pytest.mark.parametrize("extra", [("m_avg", {"time": "average"}, 0), ("m_first", {"kind": "first"}, 0),
                                   ("m_nansum", {}, math.nan), ("m_mixed", {"time": "highest"}, 0)])
def test_order_dependent_rules_take_the_chain(extra):
    measures = FUSABLE + [extra]
    fused, plain = _cube(FusedOracleStore, measures), _cube(OracleStore, measures)
    FusedOracleStore.calls = 0
    for op in OPS:
        _same(op(fused), op(plain))
    assert FusedOracleStore.calls == 0
