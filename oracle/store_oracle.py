"""ORACLE — test infrastructure, not product code.

CPU restatement (pure Python, small cases only) of the reference's per-measure
store, /root/reference/src/store/in-memory.js:7-431.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product package never does.

The reference keeps cells in a JavaScript ``Map<int, double>``; iteration order
is insertion order, ``set`` on an existing key keeps its slot and ``delete`` +
``set`` appends.  A Python ``dict`` has exactly those semantics, and Python
floats are IEEE doubles like JS numbers, so each method below follows the
reference loop statement by statement (file:line cited per method).

Pinned against the reference's own test expectations, transcribed in
tests/golden/reference_kats.json (the reference cannot be executed in this
image: no Node.js)."""
from __future__ import annotations

import math

_TYPE_SIZE = {"int32": 4, "uint32": 4, "float32": 4, "float64": 8}


def _js_max(a, b):
    """Math.max: NaN-propagating, max(-0, +0) = +0."""
    if a != a or b != b:
        return math.nan
    if a == 0 and b == 0:
        return a if math.copysign(1, a) > 0 else b
    return a if a > b else b


def _js_min(a, b):
    if a != a or b != b:
        return math.nan
    if a == 0 and b == 0:
        return a if math.copysign(1, a) < 0 else b
    return a if a < b else b


def _js_mod(a, b):
    """JS % : sign of the dividend (C fmod)."""
    if b == 0 or a != a or b != b or math.isinf(a):
        return math.nan
    return math.fmod(a, b)


_AGGREGATIONS = {  # in-memory.js:282-290
    "sum": lambda a, b: a + b,
    "average": lambda a, b: a + b,
    "highest": _js_max,
    "lowest": _js_min,
    "first": lambda a, _b: a,
    "last": lambda _a, b: b,
    "product": lambda a, b: a * b,
}


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        return math.nan if a == 0 or a != a else math.copysign(math.inf, a)


class OracleStore:
    # ---- in-memory.js:48-64 -------------------------------------------
    def __init__(self, size, type="float32", defaultValue=math.nan, dataMap=None):
        self._size = size
        self._type = type
        if not (defaultValue != defaultValue) and defaultValue != 0:
            raise ValueError("Invalid default value, only NaN and 0 are supported")
        if type not in _TYPE_SIZE:
            raise ValueError("Invalid type")
        self._defaultValue = defaultValue
        self._dataMap = dict(dataMap or {})

    # ---- in-memory.js:8-46 ----------------------------------------------
    @property
    def byteLength(self):
        return self._size * _TYPE_SIZE.get(self._type, 1)

    @property
    def size(self):
        return self._size

    @property
    def total(self):
        total = 0.0
        for value in self._dataMap.values():
            total += value
        return total

    @property
    def data(self):
        result = [self._defaultValue] * self._size
        for index, value in self._dataMap.items():
            result[index] = value
        return result

    @data.setter
    def data(self, values):
        if self._size != len(values):
            raise ValueError(f"value length is invalid: {self._size} !== {len(values)}")
        for i in range(self._size):
            self.setValue(i, values[i])

    @property
    def status(self):
        """README.md:698-721 semantics restricted to what this fork can tell:
        0x2 when the cell is set, 0x1 when it is not."""
        return [0x2 if i in self._dataMap else 0x1 for i in range(self._size)]

    def serialize(self):  # in-memory.js:75-101 (keys in Map order)
        from olap_in_memory_b200.serialization import store_to_buffer

        return store_to_buffer(self._size, self._type, self._defaultValue, list(self._dataMap.keys()),
                               list(self._dataMap.values()))

    @classmethod
    def deserialize(cls, buffer, size=None):  # in-memory.js:103-116: no presence filter
        from olap_in_memory_b200.serialization import store_from_buffer

        wire_size, type, default, keys, values = store_from_buffer(buffer)
        store = cls(wire_size if size is None else size, type, default)
        store._dataMap = dict(zip(keys.tolist(), values.tolist()))
        return store

    def clone(self):  # in-memory.js:66-73
        return OracleStore(self._size, self._type, self._defaultValue, self._dataMap)

    # ---- in-memory.js:118-137 ---------------------------------------------
    def getValue(self, index):
        return self._dataMap.get(index, self._defaultValue)

    def setValue(self, index, value):
        d = self._defaultValue
        if value is not None and not (value == d) and not (d != d and value != value):
            self._dataMap[index] = float(value)
        else:
            self._dataMap.pop(index, None)

    def fill(self, value):
        for i in range(self._size):
            self.setValue(i, value)

    # ---- in-memory.js:139-176 ---------------------------------------------
    def load(self, otherStore, myDimensions, hisDimensions):
        n = len(myDimensions)
        his_len = [d.numItems for d in hisDimensions]
        my_len = [d.numItems for d in myDimensions]
        his_to_mine = []
        for i, his in enumerate(hisDimensions):
            mine = myDimensions[i].getItemsToIdx()
            his_to_mine.append([mine.get(item) for item in his.getItems()])
        self.load_lowered(otherStore, my_len, his_len, his_to_mine)

    def load_lowered(self, otherStore, my_len, his_len, his_to_mine):
        n = len(my_len)
        coord = [0] * n
        for other_idx in range(otherStore._size):
            rest = other_idx
            for i in range(n - 1, -1, -1):
                coord[i] = rest % his_len[i]
                rest //= his_len[i]
            my_idx = 0
            for i in range(n):
                offset = his_to_mine[i][coord[i]]
                if offset is None or offset < 0:
                    # item unknown to me: the reference computes `idx + undefined` = NaN and
                    # writes to the phantom Map key NaN (never read back by data/getValue).
                    # Declared divergence: the cell is dropped ("load").
                    my_idx = None
                    break
                my_idx = my_idx * my_len[i] + offset
            if my_idx is not None:
                self.setValue(my_idx, otherStore.getValue(other_idx))

    # ---- in-memory.js:178-211 ---------------------------------------------
    def reorder(self, oldDimensions, newDimensions):
        new_to_old = [oldDimensions.index(d) for d in newDimensions]
        return self.reorder_lowered([d.numItems for d in oldDimensions], new_to_old)

    def reorder_lowered(self, old_len, new_to_old):
        out = OracleStore(self._size, self._type, self._defaultValue)
        n = len(old_len)
        new_len = [old_len[k] for k in new_to_old]
        coord = [0] * n
        for old_idx, value in self._dataMap.items():
            rest = old_idx
            for i in range(n - 1, -1, -1):
                coord[i] = rest % old_len[i]
                rest //= old_len[i]
            new_idx = 0
            for i in range(n):
                new_idx = new_idx * new_len[i] + coord[new_to_old[i]]
            out.setValue(new_idx, value)
        return out

    # ---- in-memory.js:213-263 ---------------------------------------------
    def dice(self, oldDimensions, newDimensions):
        new_to_old = []
        for i, new_dim in enumerate(newDimensions):
            old_idx = oldDimensions[i].getItemsToIdx()
            new_to_old.append([old_idx[item] for item in new_dim.getItems()])
        return self.dice_lowered([d.numItems for d in oldDimensions], new_to_old)

    def dice_lowered(self, old_len, new_to_old):
        n = len(old_len)
        new_len = [len(m) for m in new_to_old]
        new_size = math.prod(new_len)
        old_to_new = [{old: new for new, old in enumerate(m)} for m in new_to_old]
        out = OracleStore(new_size, self._type, self._defaultValue)
        coord = [0] * n
        for old_idx, value in self._dataMap.items():
            rest = old_idx
            kept = True
            for i in range(n - 1, -1, -1):
                new_coord = old_to_new[i].get(rest % old_len[i])
                if new_coord is None:
                    kept = False
                    break
                coord[i] = new_coord
                rest //= old_len[i]
            if not kept:
                continue
            new_idx = 0
            for i in range(n):
                new_idx = new_idx * new_len[i] + coord[i]
            out.setValue(new_idx, value)
        return out

    # ---- in-memory.js:265-334 ---------------------------------------------
    def drillUp(self, oldDimensions, newDimensions, method="sum"):
        maps = [
            oldDimensions[i].getGroupIndexFromRootIndexMap(new_dim.rootAttribute)
            for i, new_dim in enumerate(newDimensions)
        ]
        return self.drillUp_lowered(
            [d.numItems for d in oldDimensions], [d.numItems for d in newDimensions], maps, method
        )

    def drillUp_lowered(self, old_len, new_len, maps, method="sum"):
        method = method or "sum"
        n = len(old_len)
        new_size = math.prod(new_len)
        out = OracleStore(new_size, self._type, self._defaultValue)
        aggregate = _AGGREGATIONS.get(method)
        if aggregate is None:
            raise ValueError(f"Unsupported aggregation method: {method}")
        contributions = {}
        coord = [0] * n
        for old_idx, value in self._dataMap.items():
            rest = old_idx
            for i in range(n - 1, -1, -1):
                coord[i] = rest % old_len[i]
                rest //= old_len[i]
            new_idx = 0
            for i in range(n):
                new_idx = new_idx * new_len[i] + int(maps[i][coord[i]])
            if new_idx not in out._dataMap:
                out.setValue(new_idx, value)
            else:
                out.setValue(new_idx, aggregate(out.getValue(new_idx), value))
            contributions[new_idx] = (contributions.get(new_idx, 0) + 1) & 0xFFFF  # Uint16Array, :278
        if method == "average":
            for new_idx in range(new_size):
                count = contributions.get(new_idx, 0)
                if count:
                    out.setValue(new_idx, out.getValue(new_idx) / count)
        return out

    # ---- in-memory.js:336-430 ---------------------------------------------
    def drillDown(self, oldDimensions, newDimensions, method="sum", distributions=None):
        maps = [
            newDimensions[i].getGroupIndexFromRootIndexMap(old_dim.rootAttribute)
            for i, old_dim in enumerate(oldDimensions)
        ]
        return self.drillDown_lowered(
            [d.numItems for d in oldDimensions],
            [d.numItems for d in newDimensions],
            maps,
            method,
            distributions,
        )

    def drillDown_lowered(self, old_len, new_len, maps, method="sum", distributions=None):
        method = method or "sum"
        use_rounding = self._type in ("int32", "uint32")
        old_size = self._size
        new_size = math.prod(new_len)
        n = len(new_len)
        contributions_ids = {}
        contributions_total = {}
        idx_new_old = [0] * new_size
        coord = [0] * n
        for new_idx in range(new_size):
            rest = new_idx
            for i in range(n - 1, -1, -1):
                coord[i] = rest % new_len[i]
                rest //= new_len[i]
            old_idx = 0
            for j in range(n):
                old_idx = old_idx * old_len[j] + int(maps[j][coord[j]])
            idx_new_old[new_idx] = old_idx
            contributions_total[old_idx] = contributions_total.get(old_idx, 0) + 1

        out = OracleStore(new_size, self._type, self._defaultValue)
        for new_idx in range(new_size):
            old_idx = idx_new_old[new_idx]
            old_value = self._dataMap.get(old_idx)
            if old_value is None or old_value == 0 or old_value != old_value:  # `if (!oldValue) continue`
                continue
            num = contributions_total[old_idx]
            if distributions is not None:
                added = new_size / old_size
                shared = len(distributions) / added
                dist_index = int(math.floor(new_idx / (new_size / shared)) * added + (new_idx % added))
                if dist_index >= len(distributions) or distributions[dist_index] is None:
                    raise ValueError(f"distribution missing for index {dist_index}")
                out.setValue(new_idx, old_value * distributions[dist_index])
            elif method == "sum":
                if use_rounding:
                    value = math.floor(old_value / num)
                    remainder = _js_mod(old_value, num)
                    k = contributions_ids.get(old_idx, 0)
                    step = remainder / num
                    last_is_same = math.floor(k * step) == math.floor((k - 1) * step)
                    out.setValue(new_idx, float(value if last_is_same else value + 1))
                else:
                    out.setValue(new_idx, old_value / num)
            else:
                out.setValue(new_idx, old_value)
            contributions_ids[old_idx] = contributions_ids.get(old_idx, 0) + 1
        return out

    # ---- cube.js:326-363 (computed measures) --------------------------------
    @staticmethod
    def evaluate(expression, cell_names, stores, totals, size):
        params = dict(totals)
        result = [0.0] * size
        for i in range(size):
            for name, store in zip(cell_names, stores):
                params[name] = store.getValue(i)
            result[i] = expression.evaluate(params)
        return result
