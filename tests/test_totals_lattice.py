"""getNestedObject(s)(withTotals): the lattice walk must give exactly what the reference's
per-mask chains give (cube.js:429-439, 454-462) with 2^D - 1 drillUps instead of D*2^(D-1)."""
import math

import numpy as np

from olap_in_memory_b200 import Cube, GenericDimension, TimeDimension
from olap_in_memory_b200.cube import _deep_merge
from oracle.store_oracle import OracleStore


def _cube():
    cube = Cube([TimeDimension("time", "month", "2010-01", "2010-06"),
                 GenericDimension("location", "city", ["paris", "toledo", "tokyo"]),
                 GenericDimension("kind", "k", ["a", "b"])], OracleStore)
    rng = np.random.default_rng(5)
    for name, rules, default in (("ms", {}, 0), ("ma", {"time": "average", "location": "highest"}, math.nan),
                                 ("mf", {"time": "first", "kind": "last", "location": "lowest"}, 0)):
        cube.createStoredMeasure(name, rules, "float32", default)
        data = rng.integers(-5, 20, cube.storeSize).astype(float)
        data[rng.random(cube.storeSize) < 0.3] = default
        cube.setData(name, data.tolist())
    cube.createComputedMeasure("mc", "ms + mf")
    return cube


def _reference_totals(cube, measure_ids):
    result = {}
    for j in range(2 ** len(cube.dimensions)):
        sub = cube
        for i, dim in enumerate(cube.dimensions):
            if j & (1 << i):
                sub = sub.drillUp(dim.id, "all")
        _deep_merge(result, {m: sub.getNestedObject(m, False) for m in measure_ids})
    return result


def _same(a, b):
    if isinstance(a, dict):
        return isinstance(b, dict) and list(a) == list(b) and all(_same(a[k], b[k]) for k in a)
    return (a == b) or (a != a and b != b)


def test_lattice_equals_per_mask_chains():
    cube = _cube()
    ids = ["ms", "ma", "mf", "mc"]
    assert _same(cube.getNestedObjects(ids, True), _reference_totals(cube, ids))
    for m in ids:
        assert _same(cube.getNestedObject(m, True), _reference_totals(cube, [m])[m])


def test_lattice_visits_every_mask_once():
    cube = _cube()
    masks = [mask for mask, _ in cube._totals_lattice()]
    assert sorted(masks) == list(range(8))
    calls = []
    orig = Cube.drillUp
    try:
        Cube.drillUp = lambda self, d, a: (calls.append(d), orig(self, d, a))[1]
        cube.getNestedObjects(["ms"], True)
    finally:
        Cube.drillUp = orig
    assert len(calls) == 2 ** 3 - 1
