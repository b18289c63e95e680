"""Wire format of serialize()/deserialize() against bytes written out by
hand from /root/reference/src/serialization.js:27-89, plus the reference's own round-trip
test (test/cube-serialize.js:7-27)."""
import math
import struct

import numpy as np
import pytest

from olap_in_memory_b200.dimension import DimensionFactory, GenericDimension, TimeDimension
from olap_in_memory_b200.serialization import fromBuffer, store_from_buffer, store_to_buffer, toBuffer, undefined


def u32(*v):
    return struct.pack(f"<{len(v)}I", *v)


def f32(v):
    return struct.pack("<f", v)


def test_scalars_byte_for_byte():
    assert toBuffer(None) == u32(6)  # serialization.js:27-29
    assert toBuffer(undefined) == u32(0)  # 30-32: Uint32Array.set([undefined]) stores 0
    assert toBuffer(32) == u32(7) + f32(32.0)  # 71-74
    assert toBuffer(0.1) == u32(7) + f32(np.float32(0.1))  # numbers travel as Float32
    assert toBuffer(math.nan)[:4] == u32(7) and math.isnan(struct.unpack("<f", toBuffer(math.nan)[4:])[0])
    assert toBuffer(True) == u32(8) + f32(1.0) and toBuffer(False) == u32(8) + f32(0.0)  # 75-78


def test_buffers_and_typed_arrays_byte_for_byte():
    assert toBuffer(b"\x01\x02\x03") == u32(1, 3) + b"\x01\x02\x03\x00"  # 33-39: padded to 4
    assert toBuffer(b"") == u32(1, 0)
    # 40-48: [TYPED_ARRAY][index in TypedArraySubClasses (1-13)][ARRAY_BUFFER]
    assert toBuffer(np.array([255], dtype=np.int32)) == u32(2, 5) + u32(1, 4) + u32(255)
    assert toBuffer(np.array([1, 2], dtype=np.uint32)) == u32(2, 6) + u32(1, 8) + u32(1, 2)
    assert toBuffer(np.array([666], dtype=np.float32)) == u32(2, 7) + u32(1, 4) + f32(666.0)
    assert toBuffer(np.array([1.5], dtype=np.float64)) == u32(2, 8) + u32(1, 8) + struct.pack("<d", 1.5)
    assert toBuffer(np.array([7], dtype=np.uint8)) == u32(2, 1) + u32(1, 1) + b"\x07\0\0\0"


def test_strings_arrays_objects_byte_for_byte():
    # 65-70: [STRING] + TypedArray(Uint8Array) of the UTF-8 bytes
    ab = u32(4) + u32(2, 1) + u32(1, 2) + b"ab\0\0"
    assert toBuffer("ab") == ab
    assert toBuffer("é") == u32(4) + u32(2, 1) + u32(1, 2) + "é".encode() + b"\0\0"
    # 49-64: [ARRAY][n] then n x ([byteLength][item])
    assert toBuffer([None, 1]) == u32(3, 2) + u32(4) + u32(6) + u32(8) + u32(7) + f32(1.0)
    assert toBuffer([]) == u32(3, 0)
    # 79-87: [OBJECT] + Array of [key, ArrayBuffer(toBuffer(value))]
    value = u32(1, 8) + u32(7) + f32(1.0)
    entry = u32(3, 2) + u32(len(ab)) + ab + u32(len(value)) + value
    assert toBuffer({"ab": 1}) == u32(5) + u32(3, 1) + u32(len(entry)) + entry
    # Object.entries: array-index keys first, ascending; then insertion order
    assert list(fromBuffer(toBuffer({"b": 1, "10": 2, "a": 3, "2": 4, "01": 5})).keys()) == ["2", "10", "b", "a", "01"]
    assert list(fromBuffer(toBuffer({"b": 1, "\u00b9": 2, "7": 3})).keys()) == ["7", "b", "\u00b9"]  # only ASCII digits form an index


def test_reference_round_trip():  # test/cube-serialize.js:7-27
    obj = [math.nan, 32, np.array([255], dtype=np.int32), "totot", np.array([666], dtype=np.float32),
           {"toto": {"tata": np.array([666], dtype=np.float32)}}, None]
    new = fromBuffer(toBuffer(obj))
    assert math.isnan(new[0]) and new[1] == 32 and new[3] == "totot" and new[6] is None
    assert new[2].dtype == np.int32 and new[2].tolist() == [255]
    assert new[4].dtype == np.float32 and new[4].tolist() == [666]
    assert list(new[5]) == ["toto"] and new[5]["toto"]["tata"].tolist() == [666]
    assert fromBuffer(u32(0)) is undefined
    assert toBuffer(new) == toBuffer(obj)


def test_store_record():  # in-memory.js:75-116
    buf = store_to_buffer(6, "float32", math.nan, [4, 1], [2.5, -1.0])
    data = fromBuffer(buf)
    assert list(data) == ["size", "type", "defaultValue", "indexes", "dataBuffer"]
    assert data["size"] == 6 and data["type"] == "float32" and math.isnan(data["defaultValue"])
    assert data["indexes"].dtype == np.uint32 and data["indexes"].tolist() == [4, 1]  # Map order kept
    assert data["dataBuffer"].dtype == np.float32 and data["dataBuffer"].tolist() == [2.5, -1.0]
    size, type_, default, keys, values = store_from_buffer(buf)
    assert (size, type_) == (6, "float32") and keys.tolist() == [4, 1] and values.tolist() == [2.5, -1.0]
    # typed payloads: `new Int32Array(doubles)` truncates and wraps, `new Uint32Array` wraps negatives
    assert fromBuffer(store_to_buffer(3, "int32", 0, [0, 1, 2], [1.9, -2.9, 2.0**31]))["dataBuffer"].tolist() == [1, -2, -(2**31)]
    assert fromBuffer(store_to_buffer(3, "uint32", 0, [0, 1, 2], [1.9, -1.0, math.inf]))["dataBuffer"].tolist() == [1, 2**32 - 1, 0]
    assert fromBuffer(store_to_buffer(1, "float64", 0, [0], [0.1]))["dataBuffer"].dtype == np.float64
    # `size` is what a Float32 keeps of it (serialization.js:71-74)
    assert store_from_buffer(store_to_buffer(16777217, "float32", 0, [], []))[0] == 16777216
    with pytest.raises(OverflowError):
        store_to_buffer(2**33, "float32", 0, [2**32], [1.0])


def test_dimension_records():  # generic.js:47-72, time.js:28-47, factory.js:5-15
    location = GenericDimension("location", "city", ["paris", "toledo", "tokyo"], "Location", {"paris": "Paris"})
    location.addAttribute("city", "continent", {"paris": "europe", "toledo": "europe", "tokyo": "asia"})
    data = fromBuffer(location.serialize())
    assert list(data) == ["id", "label", "rootAttribute", "rootItems", "attributeItems", "attributeLabels", "attributeMappings"]
    assert data["rootItems"] == ["paris", "toledo", "tokyo"]
    assert data["attributeMappings"]["continent"].dtype == np.uint32
    assert data["attributeMappings"]["continent"].tolist() == [0, 0, 1]
    new = DimensionFactory.deserialize(location.serialize())
    assert isinstance(new, GenericDimension) and new.id == "location" and new.label == "Location"
    assert new.attributes == location.attributes and new.getItems("continent") == ["europe", "asia"]
    assert new.getGroupIndexFromRootIndexMap("continent").tolist() == [0, 0, 1]
    assert new.getEntries() == location.getEntries()
    assert new.serialize() == location.serialize()

    time = TimeDimension("time", "month", "2010-01", "2011-01", "Time")
    data = fromBuffer(time.serialize())
    assert data == {"id": "time", "label": "Time", "rootAttribute": "month", "start": "2010-01-01", "end": "2011-01-31"}
    new = DimensionFactory.deserialize(time.serialize())
    assert isinstance(new, TimeDimension) and new.getItems() == time.getItems() and new.serialize() == time.serialize()


# ---- property tests: every value the format can carry survives the wire, records stay 4-byte aligned ----
from hypothesis import given, settings, strategies as st  # noqa: E402

_f32 = st.floats(width=32, allow_nan=False)
_leaf = st.one_of(
    st.none(), st.booleans(), _f32, st.text(max_size=12), st.binary(max_size=9),
    st.lists(st.integers(-2**31, 2**31 - 1), max_size=5).map(lambda v: np.asarray(v, dtype=np.int32)),
    st.lists(_f32, max_size=5).map(lambda v: np.asarray(v, dtype=np.float32)),
    st.lists(st.floats(allow_nan=False), max_size=5).map(lambda v: np.asarray(v, dtype=np.float64)),
)
_tree = st.recursive(_leaf, lambda kids: st.one_of(st.lists(kids, max_size=4), st.dictionaries(st.text(max_size=6), kids, max_size=4)),
                     max_leaves=12)


def _same(a, b):
    if isinstance(a, np.ndarray):
        return isinstance(b, np.ndarray) and a.dtype == b.dtype and a.tolist() == b.tolist()
    if isinstance(a, dict):
        return isinstance(b, dict) and set(a) == set(b) and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, list):
        return isinstance(b, list) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return type(a) is type(b) and a == b


@settings(max_examples=300, deadline=None, derandomize=True)
@given(_tree)
def test_any_tree_round_trips(tree):
    buf = toBuffer(tree)
    assert len(buf) % 4 == 0  # every record keeps the Uint32Array views of fromBuffer aligned (serialization.js:91-92)
    back = fromBuffer(buf)
    assert _same(tree, back), (tree, back)
    # keys come back in JavaScript's enumeration order; from there on the bytes are a fixed point
    assert toBuffer(fromBuffer(toBuffer(back))) == toBuffer(back)


@settings(max_examples=150, deadline=None, derandomize=True)
@given(st.integers(0, 200).flatmap(lambda n: st.tuples(
    st.just(n), st.sampled_from(["float32", "int32", "uint32", "float64"]), st.sampled_from([0.0, math.nan]),
    st.lists(st.integers(0, max(n - 1, 0)), unique=True, max_size=n), st.randoms(use_true_random=False))))
def test_oracle_store_round_trips_in_map_order(args):
    from oracle.store_oracle import OracleStore

    n, type_, default, keys, rnd = args
    store = OracleStore(n, type_, default)
    for k in keys:  # insertion order = Map order, not ascending
        store.setValue(k, float(rnd.randint(1, 1000)))
    back = OracleStore.deserialize(store.serialize())
    assert (back._size, back._type) == (n, type_)
    assert list(back._dataMap.items()) == list(store._dataMap.items())
    assert (back._defaultValue != back._defaultValue) == (default != default)
    assert back.serialize() == store.serialize()


def _golden_b64():
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "serialized_test_cube.b64")) as f:
        return f.read().strip()


def _check_golden_cube(store_cls):
    from olap_in_memory_b200 import Cube

    cube = Cube.deserializeFromBase64String(_golden_b64(), store_cls)
    assert cube.dimensionIds == ["location", "period"] and cube.storedMeasureIds == ["antennas", "routers"]
    assert cube.getDimension("location").attributes == ["all", "city", "country", "continent", "citySize"]
    assert cube.getNestedArray("antennas") == [[1, 2], [4, 0], [16, 32]]
    assert cube.getNestedArray("routers") == [[3, 2], [4, 9], [16, 32]]
    assert cube.drillUp("location", "continent").getNestedArray("antennas") == [[5, 2], [16, 32]]
    assert cube.computedMeasureIds == ["router_by_antennas"]
    # ascending keys = the Map order of this fixture: the device store and the oracles emit the same bytes
    assert cube.serializeToBase64String() == _golden_b64()


def test_golden_serialized_cube_on_the_oracles():  # tests/golden/make_serialized.py
    from oracle.c_oracle import COracleStore
    from oracle.store_oracle import OracleStore

    _check_golden_cube(OracleStore)
    _check_golden_cube(COracleStore)


@pytest.mark.gpu
def test_golden_serialized_cube_on_the_device():
    from olap_in_memory_b200.store import GpuStore

    _check_golden_cube(GpuStore)
