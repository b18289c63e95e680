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
    def drillUp_lowered_batch(stores, old_len, new_len, maps, methods):
        FusedOracleStore.calls += 1
        return [OracleStore.drillUp_lowered(s, old_len, new_len, maps, m) for s, m in zip(stores, methods)]


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
    for op in OPS:
        _same(op(fused), op(plain))  # fused where the removed dimensions allow it, chained otherwise
    FusedOracleStore.calls = 0
    _same(fused.collapse(), plain.collapse())  # removes every dimension: the odd rule is always involved
    assert FusedOracleStore.calls == 0


@pytest.mark.gpu
@pytest.mark.parametrize("op", range(len(OPS)))
def test_gpu_fused_equals_reference_chain(op):
    """GpuStore opts in (FUSED_ROLLUPS): the fused device path against the oracle's chain."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200.store import GpuStore

    N.init(0)
    assert GpuStore.FUSED_ROLLUPS
    gpu, plain = _cube(GpuStore, FUSABLE), _cube(OracleStore, FUSABLE)
    launches = N.lib().olap_kernel_launches()
    got, want = OPS[op](gpu), OPS[op](plain)
    assert N.lib().olap_kernel_launches() > launches
    assert got.dimensionIds == want.dimensionIds and got.storedMeasuresRules == want.storedMeasuresRules
    for m in got.storedMeasureIds + got.computedMeasureIds:
        x = np.asarray(got.getData(m), np.float32)
        y = np.asarray(want.getData(m), float).astype(np.float32)
        assert np.array_equal(x, y, equal_nan=True), m


@pytest.mark.gpu
def test_gpu_collapse_of_a_large_cube_is_one_pass():
    """collapse() of 8^7 = 2 097 152 cells x 2 measures: one store pass on the long-row kernel."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200.store import GpuStore

    N.init(0)
    dims = [GenericDimension(f"d{i}", "root", [f"d{i}-{j}" for j in range(8)]) for i in range(7)]
    cube = Cube(dims, GpuStore)
    rng = np.random.default_rng(2)
    data = {}
    for name, rule in (("m_sum", None), ("m_hi", "highest")):
        cube.createStoredMeasure(name, {d.id: rule for d in dims} if rule else {}, "float32", 0)
        data[name] = rng.integers(-100, 1000, cube.storeSize).astype(np.float32)
        data[name][rng.random(cube.storeSize) < 0.3] = 0.0
        cube.setData(name, data[name])
    launches = N.lib().olap_kernel_launches()
    out = cube.collapse()
    assert N.lib().olap_last_op_path().decode() == "drillup/long"
    assert N.lib().olap_kernel_launches() - launches <= 2  # rollup + merge pass
    assert out.dimensionIds == []
    assert out.getData("m_sum") == [float(np.float32(data["m_sum"].astype(np.float64).sum()))]
    assert out.getData("m_hi") == [float(data["m_hi"][data["m_hi"] != 0].max())]


@pytest.mark.gpu
def test_gpu_implied_maps_equal_explicit_maps():
    """include/olap_gpu.h: maps[d] == NULL means 'unchanged' or 'everything to the one new item'."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200.store import GpuStore as G

    N.init(0)
    rng = np.random.default_rng(4)
    dims = [6, 50, 7]
    s = G(int(np.prod(dims)), "float32", 0)
    s.set_data_f32(rng.integers(-3, 9, s.size).astype(np.float32))
    ident = [np.arange(n, dtype=np.int32) for n in dims]
    for new_len, explicit, implied in (
        ([6, 1, 7], [ident[0], np.zeros(50, np.int32), ident[2]], [None, None, None]),          # one axis -> all
        ([1, 50, 1], [np.zeros(6, np.int32), ident[1], np.zeros(7, np.int32)], [None, None, None]),  # two axes: generic kernel
        ([6, 50, 7], ident, [None, None, None]),                                                   # plain copy
    ):
        for method in ("sum", "first", "average"):
            a = G.drillUp_lowered([s], dims, new_len, explicit, [method])[0].data_f32()
            b = G.drillUp_lowered([s], dims, new_len, implied, [method])[0].data_f32()
            assert np.array_equal(a, b, equal_nan=True), (new_len, method)
    with pytest.raises(Exception):
        G.drillUp_lowered([s], dims, [6, 5, 7], [None, None, None], ["sum"])
