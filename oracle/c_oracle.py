"""ORACLE — test infrastructure, not product code.

ctypes wrapper of oracle/olap_oracle.c (the plain-C restatement of
/root/reference/src/store/in-memory.js).  Same lowered interface as
oracle/store_oracle.OracleStore (`*_lowered` methods taking dimension lengths and
int32 maps), so the parity tests feed identical arguments to the C oracle, the Python
oracle and the CUDA store.  Used by tests/ and by bench.py's cpu_baseline leg only."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle.so")
_TYPES = {"int32": 0, "uint32": 1, "float32": 2, "float64": 3}
_METHODS = {"sum": 0, "average": 1, "highest": 2, "lowest": 3, "first": 4, "last": 5, "product": 6}
_lib = None
_p_i32 = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(os.path.join(_HERE, "olap_oracle.c")):
            subprocess.run(["make", "-s"], cwd=_HERE, check=True)
        L = C.CDLL(_PATH)
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        sig = {
            "ostore_new": (vp, [i64, C.c_int, C.c_int]),
            "ostore_free": (None, [vp]),
            "ostore_clone": (vp, [vp]),
            "ostore_get_value": (dbl, [vp, i64]),
            "ostore_set_value": (None, [vp, i64, dbl]),
            "ostore_set_data": (C.c_int, [vp, vp, i64]),
            "ostore_set_data_f32": (C.c_int, [vp, vp, i64]),
            "ostore_get_data": (None, [vp, vp]),
            "ostore_fill": (None, [vp, dbl]),
            "ostore_total": (dbl, [vp]),
            "ostore_size": (i64, [vp]),
            "ostore_count": (i64, [vp]),
            "ostore_entries": (None, [vp, vp, vp]),
            "ostore_reorder": (vp, [vp, C.c_int, vp, vp]),
            "ostore_dice": (vp, [vp, C.c_int, vp, vp, vp]),
            "ostore_drill_up": (vp, [vp, C.c_int, vp, vp, vp, C.c_int]),
            "ostore_drill_down": (vp, [vp, C.c_int, vp, vp, vp, C.c_int, vp, i64, C.POINTER(i64)]),
            "ostore_load": (None, [vp, vp, C.c_int, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _i64(values):
    """ctypes int64 array (passed by reference, alive for the duration of the call)."""
    return (C.c_int64 * max(1, len(values)))(*[int(v) for v in values])


def _maps(maps):
    keep = [np.ascontiguousarray(np.asarray(m, dtype=np.int32)) for m in maps]
    ptrs = (_p_i32 * max(1, len(keep)))(*[a.ctypes.data_as(_p_i32) for a in keep])
    return keep, ptrs


class COracleStore:
    def __init__(self, size, type="float32", defaultValue=math.nan, _handle=None):
        self._type = type
        self._defaultValue = defaultValue
        nan_default = defaultValue != defaultValue
        if not nan_default and defaultValue != 0:
            raise ValueError("Invalid default value, only NaN and 0 are supported")
        self._h = _handle if _handle is not None else lib().ostore_new(int(size), _TYPES[type], int(nan_default))
        self._size = lib().ostore_size(self._h)

    def _wrap(self, handle):
        return COracleStore(0, self._type, self._defaultValue, _handle=handle)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ostore_free(h)

    @property
    def size(self):
        return self._size

    @property
    def total(self):
        return lib().ostore_total(self._h)

    @property
    def count(self):
        return lib().ostore_count(self._h)

    def data_f64(self):
        out = np.empty(self._size, dtype=np.float64)
        lib().ostore_get_data(self._h, out.ctypes.data)
        return out

    @property
    def data(self):
        return self.data_f64().tolist()

    @data.setter
    def data(self, values):
        d = float(self._defaultValue)  # undefined / null unset the cell (in-memory.js:124-125)
        arr = np.fromiter((d if v is None else v for v in values), dtype=np.float64, count=len(values))
        if lib().ostore_set_data(self._h, arr.ctypes.data, arr.size) != 0:
            raise ValueError(f"value length is invalid: {self._size} !== {arr.size}")

    def set_data_f32(self, values):
        arr = np.ascontiguousarray(values, dtype=np.float32)
        if lib().ostore_set_data_f32(self._h, arr.ctypes.data, arr.size) != 0:
            raise ValueError(f"value length is invalid: {self._size} !== {arr.size}")

    def entries(self):
        n = self.count
        keys = np.empty(n, dtype=np.int64)
        vals = np.empty(n, dtype=np.float64)
        lib().ostore_entries(self._h, keys.ctypes.data, vals.ctypes.data)
        return keys, vals

    def getValue(self, index):
        return lib().ostore_get_value(self._h, int(index))

    def setValue(self, index, value):
        lib().ostore_set_value(self._h, int(index), float(self._defaultValue if value is None else value))

    def fill(self, value):
        lib().ostore_fill(self._h, float(value))

    def serialize(self):  # in-memory.js:75-101 (Map order)
        from olap_in_memory_b200.serialization import store_to_buffer

        keys, vals = self.entries()
        return store_to_buffer(self._size, self._type, self._defaultValue, keys, vals)

    @classmethod
    def deserialize(cls, buffer, size=None):  # in-memory.js:103-116
        from olap_in_memory_b200.serialization import store_from_buffer

        wire_size, type, default, keys, values = store_from_buffer(buffer)
        store = cls(wire_size if size is None else size, type, default)
        for k, v in zip(keys.tolist(), values.tolist()):
            store.setValue(k, v)
        return store

    def clone(self):
        return self._wrap(lib().ostore_clone(self._h))

    def reorder_lowered(self, old_len, new_to_old):
        perm = np.ascontiguousarray(new_to_old, dtype=np.int32)
        return self._wrap(lib().ostore_reorder(self._h, len(old_len), _i64(old_len), perm.ctypes.data_as(C.c_void_p)))

    def dice_lowered(self, old_len, keep):
        alive, ptrs = _maps(keep)
        new_len = _i64([len(k) for k in keep])
        return self._wrap(lib().ostore_dice(self._h, len(old_len), _i64(old_len), new_len, ptrs))

    def drillUp_lowered(self, old_len, new_len, maps, method="sum"):
        method = method or "sum"
        if method not in _METHODS:
            raise ValueError(f"Unsupported aggregation method: {method}")
        alive, ptrs = _maps(maps)
        return self._wrap(
            lib().ostore_drill_up(self._h, len(old_len), _i64(old_len), _i64(new_len), ptrs,
                                  _METHODS[method])
        )

    def drillDown_lowered(self, old_len, new_len, maps, method="sum", distributions=None):
        alive, ptrs = _maps(maps)
        err = C.c_int64(-1)
        if distributions is None:
            dist_ptr, dist_len = None, 0
        else:
            dist = np.array([math.nan if v is None else v for v in distributions], dtype=np.float64)
            dist_ptr, dist_len = dist.ctypes.data, dist.size
        h = lib().ostore_drill_down(self._h, len(old_len), _i64(old_len), _i64(new_len), ptrs,
                                    int((method or "sum") == "sum"), dist_ptr, dist_len, C.byref(err))
        if not h:
            raise ValueError(f"distribution missing for index {err.value}")
        return self._wrap(h)

    def load_lowered(self, other, my_len, his_len, his_to_mine):
        alive, ptrs = _maps([[-1 if v is None else v for v in m] for m in his_to_mine])
        lib().ostore_load(self._h, other._h, len(my_len), _i64(my_len), _i64(his_len), ptrs)

    # ---- the reference's dimension-object interface (in-memory.js:139-430), lowered like
    # ---- the device store lowers it, so that Cube(dims, store_cls=COracleStore) runs the
    # ---- reference algorithm in C (bench_cube_benchmark.py's CPU column)
    @property
    def byteLength(self):
        return self._size * (8 if self._type == "float64" else 4)

    @property
    def _dataMap(self):
        keys, vals = self.entries()
        return dict(zip(keys.tolist(), vals.tolist()))

    @property
    def status(self):
        present = set(self.entries()[0].tolist())
        return [2 if i in present else 1 for i in range(self._size)]

    def reorder(self, oldDimensions, newDimensions):
        return self.reorder_lowered([d.numItems for d in oldDimensions],
                                    [next(i for i, d in enumerate(oldDimensions) if d is nd) for nd in newDimensions])

    def dice(self, oldDimensions, newDimensions):
        keep = []
        for i, new_dim in enumerate(newDimensions):
            old_idx = oldDimensions[i].getItemsToIdx()
            keep.append([old_idx[item] for item in new_dim.getItems()])
        return self.dice_lowered([d.numItems for d in oldDimensions], keep)

    def drillUp(self, oldDimensions, newDimensions, method="sum"):
        maps = [oldDimensions[i].getGroupIndexFromRootIndexMap(nd.rootAttribute) for i, nd in enumerate(newDimensions)]
        return self.drillUp_lowered([d.numItems for d in oldDimensions], [d.numItems for d in newDimensions], maps, method)

    def drillDown(self, oldDimensions, newDimensions, method="sum", distributions=None):
        maps = [newDimensions[i].getGroupIndexFromRootIndexMap(od.rootAttribute) for i, od in enumerate(oldDimensions)]
        return self.drillDown_lowered([d.numItems for d in oldDimensions], [d.numItems for d in newDimensions], maps,
                                      method, distributions)

    def load(self, otherStore, myDimensions, hisDimensions):
        his_to_mine = []
        for i, his in enumerate(hisDimensions):
            mine = myDimensions[i].getItemsToIdx()
            his_to_mine.append([mine.get(item) for item in his.getItems()])
        self.load_lowered(otherStore, [d.numItems for d in myDimensions], [d.numItems for d in hisDimensions], his_to_mine)

    @staticmethod
    def evaluate(expression, cell_names, stores, totals, size):
        cols = [s.data_f64() for s in stores]
        params = dict(totals)
        out = [0.0] * size
        for i in range(size):
            for name, col in zip(cell_names, cols):
                params[name] = col[i]
            out[i] = expression.evaluate(params)
        return out
