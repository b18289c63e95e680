"""Seeded store-level cases in LOWERED form (dimension lengths + int32 maps), the form
that crosses the C ABI.  The same case objects are fed to the Python oracle, the C
oracle and the CUDA store, and their results compared (bit-exact where SURVEY.md
§8d says so).  Inputs are float32-representable so both sides start from identical
data."""
from __future__ import annotations

import math

import numpy as np

METHODS = ["sum", "average", "highest", "lowest", "first", "last", "product"]


def f32(x):
    return np.asarray(x, dtype=np.float32)


def make_data(rng, n, default, fill=0.7, kind="int"):
    """float32 cells; a cell is unset (holds the default) with probability 1-fill."""
    if kind == "int":
        v = rng.integers(-50, 1000, n).astype(np.float32)
    elif kind == "small":
        v = (rng.integers(1, 9, n) / 4.0).astype(np.float32)  # products stay finite
    else:
        v = (1.0 + rng.random(n) * 999.0).astype(np.float32)
    if kind != "small":
        # sprinkle special values that are *set* cells under one default or the other
        special = rng.random(n)
        if default != default:
            v[special < 0.03] = 0.0  # a stored 0 is present under a NaN default
            v[(special >= 0.03) & (special < 0.04)] = -0.0
        else:
            v[special < 0.01] = np.nan  # a stored NaN is present under a 0 default
    unset = rng.random(n) >= fill
    v[unset] = np.nan if default != default else 0.0
    return v


def random_map(rng, C, P, monotone):
    """child -> parent map covering (mostly) all parents."""
    if monotone:
        cuts = np.sort(rng.integers(0, C + 1, P - 1)) if P > 1 else np.array([], dtype=np.int64)
        m = np.zeros(C, dtype=np.int32)
        for c in cuts:
            m[c:] += 1
        return m
    m = rng.integers(0, P, C).astype(np.int32)
    return m


def identity(n):
    return np.arange(n, dtype=np.int32)


def drillup_cases(seed=0):
    rng = np.random.default_rng(seed)
    shapes = [
        # (dims, changed dim, P, monotone)
        ([37, 12], 0, 5, True),          # outer axis, I=12 (vec4)
        ([37, 13], 0, 5, False),         # outer axis, I=13 (scalar)
        ([6, 31, 8], 1, 4, True),        # mid axis
        ([6, 31, 8], 1, 7, False),
        ([50, 29], 1, 3, True),          # innermost axis (I=1)
        ([50, 29], 1, 6, False),
        ([3, 4, 5, 6], 2, 1, True),      # -> all
        ([365, 2], 0, 12, True),         # config-1 style
        ([9], 0, 1, True),               # 1-D -> all
        ([9], 0, 4, False),
        ([4, 1, 6], 1, 1, True),         # C == P == 1 (slice's removeDimension)
        ([2, 700, 4], 1, 3, True),       # long segments
        ([1200, 3], 0, 2, False),
        ([1000, 8], 0, 1, True),         # few outputs, long child lists: split kernel (mid)
        ([1000, 8], 0, 3, False),
        ([4, 2000], 1, 1, True),         # ... split inside the tile kernel
        ([3, 1500, 2], 1, 2, False),
        ([20, 66], 0, 3, True),          # inner run even but not a multiple of 4: 64-bit accesses
        ([5, 9, 70], 1, 2, False),
    ]
    for dims, d, P, mono in shapes:
        for default in (0.0, math.nan):
            for method in METHODS:
                kind = "small" if method == "product" else ("int" if rng.random() < 0.5 else "real")
                n = int(np.prod(dims))
                maps = [identity(x) for x in dims]
                maps[d] = random_map(rng, dims[d], P, mono)
                new_len = list(dims)
                new_len[d] = P
                yield dict(op="drillUp", old_len=list(dims), new_len=new_len, maps=maps, method=method,
                           default=default, data=make_data(rng, n, default, fill=rng.choice([1.0, 0.6, 0.15]), kind=kind),
                           type="float32")
    # several dimensions at once (store API generality, in-memory.js:270-274)
    for default in (0.0, math.nan):
        for method in METHODS:
            dims = [5, 6, 7]
            maps = [random_map(rng, 5, 2, False), identity(6), random_map(rng, 7, 3, True)]
            yield dict(op="drillUp", old_len=dims, new_len=[2, 6, 3], maps=maps, method=method, default=default,
                       data=make_data(rng, 210, default, fill=0.5, kind="small" if method == "product" else "int"),
                       type="float32")
    # a parent without any child, empty inputs
    yield dict(op="drillUp", old_len=[4, 3], new_len=[3, 3], maps=[np.array([0, 0, 2, 2], np.int32), identity(3)],
               method="sum", default=math.nan, data=make_data(rng, 12, math.nan, 1.0), type="float32")
    yield dict(op="drillUp", old_len=[0, 3], new_len=[1, 3], maps=[np.zeros(0, np.int32), identity(3)],
               method="sum", default=0.0, data=np.zeros(0, np.float32), type="float32")


def drillup_long_cases(seed=7):
    """Rows too long for one shared-memory tile with a short inner run and few parents:
    drillup_long_kernel (segments, ordered fold, scratch + merge kernel)."""
    rng = np.random.default_rng(seed)
    shapes = [
        ([60000], 0, 1, True),         # 1-D -> all: one row, many CTAs, merge kernel
        ([3, 50001], 1, 4, False),     # rows that start off 16-byte boundaries, random parents
        ([2, 30000, 3], 1, 5, True),   # I = 3
        ([45000, 2], 0, 3, False),     # O = 1, I = 2
        ([2, 11000, 7], 1, 2, True),   # I = 7
        ([1, 300000], 1, 200, False),  # 200 parents x 1500 children, outputs > 128: thread-per-output path
    ]
    for dims, d, P, mono in shapes:
        for default in (0.0, math.nan):
            for method in METHODS:
                kind = "small" if method == "product" else ("int" if rng.random() < 0.5 else "real")
                if method == "product" and dims[d] // P > 3000:
                    continue  # products of thousands of factors leave float range
                n = int(np.prod(dims))
                maps = [identity(x) for x in dims]
                maps[d] = random_map(rng, dims[d], P, mono)
                new_len = list(dims)
                new_len[d] = P
                data = make_data(rng, n, default, fill=rng.choice([1.0, 0.6, 0.05]), kind=kind)
                if default != default and method in ("sum", "average"):
                    # the restart corner (in-memory.js:311-318): +inf and -inf under one parent
                    data[rng.integers(0, n, 3)] = np.inf
                    data[rng.integers(0, n, 3)] = -np.inf
                yield dict(op="drillUp", old_len=list(dims), new_len=new_len, maps=maps, method=method,
                           default=default, data=data, type="float32")


def drilldown_cases(seed=1):
    rng = np.random.default_rng(seed)
    shapes = [
        ([4, 8], 0, 11, True),     # outer axis, I=8
        ([4, 7], 0, 11, True),     # scalar inner
        ([5, 3, 6], 1, 10, True),  # mid axis
        ([20, 4], 1, 13, True),    # innermost axis
        ([3], 0, 9, True),
        ([6, 1, 4], 1, 5, True),   # addDimension: CatchAll(1) -> 5 items
        ([3, 5], 1, 12, False),    # non-monotone new->old map
        ([4, 64], 0, 11, True),    # long inner run: parent-driven kernel, vec4
        ([3, 5, 32], 1, 9, False),
        ([2, 3, 130], 1, 7, True),  # long inner run, 64-bit accesses
        ([2, 3, 129], 1, 5, True),  # long inner run, scalar
    ]
    for dims, d, C, mono in shapes:
        for default in (0.0, math.nan):
            for typ in ("float32", "uint32", "int32"):
                for method in ("sum", "average"):
                    n = int(np.prod(dims))
                    P = dims[d]
                    m = random_map(rng, C, P, mono)
                    maps = [identity(x) for x in dims]
                    maps[d] = m
                    new_len = list(dims)
                    new_len[d] = C
                    data = make_data(rng, n, default, fill=0.7, kind="int")
                    if typ == "uint32":
                        data = np.abs(data)
                    yield dict(op="drillDown", old_len=list(dims), new_len=new_len, maps=maps, method=method,
                               default=default, data=data, type=typ, distributions=None)
    # distributions (in-memory.js:391-400): added dimension innermost, shared outermost
    yield dict(op="drillDown", old_len=[2, 1], new_len=[2, 3], maps=[identity(2), np.zeros(3, np.int32)],
               method="sum", default=0.0, data=f32([10, 20]), type="float32",
               distributions=[0.5, 0.3, 0.2, 0.1, 0.1, 0.8])
    # several dimensions at once
    for typ in ("float32", "int32"):
        yield dict(op="drillDown", old_len=[2, 3], new_len=[5, 7],
                   maps=[random_map(rng, 5, 2, True), random_map(rng, 7, 3, True)], method="sum", default=0.0,
                   data=f32([100, 7, 0, -7, 33, 1]), type=typ, distributions=None)


def dice_cases(seed=2):
    rng = np.random.default_rng(seed)
    shapes = [
        ([10, 16], {0: 5}),
        ([10, 15], {0: 4}),
        ([6, 9, 8], {1: 3}),
        ([12, 20], {1: 7}),
        ([5, 6, 7], {0: 2, 2: 3}),   # diceByDimensionItems
        ([8, 4], {0: 0}),             # empty result
        ([8, 4], {1: 1}),
        ([1, 9, 12], {1: 9}),
        # the innermost axis alone is diced: whole source rows staged (dice_inner_kernel), ragged last CTA
        ([5000, 10], {1: 5}),
        ([37, 41, 9], {2: 4}),
        ([3001, 7], {1: 6}),
        ([6, 5, 300], {2: 299}),
        ([40, 6, 11], {1: 1, 2: 3}),   # a sliced middle axis in front of the diced innermost one
    ]
    for dims, cut in shapes:
        for default in (0.0, math.nan):
            for shuffle in (False, True):
                keep = []
                for d, length in enumerate(dims):
                    if d in cut:
                        k = rng.choice(length, size=cut[d], replace=False).astype(np.int32)
                        keep.append(k if shuffle else np.sort(k))
                    else:
                        keep.append(identity(length))
                n = int(np.prod(dims))
                yield dict(op="dice", old_len=list(dims), keep=keep, default=default,
                           data=make_data(rng, n, default, 0.6, "int"), type="float32")


def reorder_cases(seed=3):
    rng = np.random.default_rng(seed)
    import itertools

    # big enough for several (ragged) boxes of the tiled transpose
    for dims, perm in (([130, 70], [1, 0]), ([37, 50, 3, 70], [3, 2, 1, 0]), ([37, 50, 3, 70], [1, 3, 0, 2]),
                       ([20, 30, 10, 10], [3, 2, 1, 0]), ([9, 300, 11], [2, 0, 1]), ([64, 64, 8], [0, 2, 1]),
                       ([3, 5000], [1, 0]), ([10, 10, 10, 10, 10], [4, 3, 2, 1, 0]),
                       # two disjoint 4-aligned axis groups: the pair transpose, ragged tiles included
                       ([100, 104], [1, 0]), ([332, 100], [1, 0]), ([12, 32, 32, 44], [3, 2, 1, 0]),
                       ([52, 7, 92], [2, 1, 0]), ([44, 4, 25, 8], [3, 2, 1, 0]), ([20, 12, 10, 10, 10], [4, 3, 1, 2, 0]),
                       ([72, 200, 12], [2, 0, 1]), ([8, 8, 8, 8, 8], [2, 4, 0, 3, 1]),
                       # 200-cell runs on both sides: the 2-CTA cluster tile (distributed shared memory), ragged too
                       ([24, 16, 10, 10, 10], [4, 3, 2, 1, 0]), ([28, 12, 6, 10, 10], [4, 3, 2, 1, 0]), ([400, 408], [1, 0])):
        for default in (0.0, math.nan):
            n = int(np.prod(dims))
            yield dict(op="reorder", old_len=list(dims), new_to_old=list(perm), default=default,
                       data=make_data(rng, n, default, 0.7, "int"), type="float32")
    for dims in ([2, 3], [4, 5, 6], [3, 1, 4, 8], [7, 16], [5, 4, 3, 2, 6]):
        perms = list(itertools.permutations(range(len(dims))))
        if len(perms) > 12:
            perms = [perms[i] for i in rng.choice(len(perms), 12, replace=False)]
        for perm in perms:
            for default in (0.0, math.nan):
                n = int(np.prod(dims))
                yield dict(op="reorder", old_len=list(dims), new_to_old=list(perm), default=default,
                           data=make_data(rng, n, default, 0.7, "int"), type="float32")


def load_cases(seed=4):
    rng = np.random.default_rng(seed)
    for my_len, his_len in (([4, 6], [2, 5]), ([3, 3, 3], [3, 2, 4]), ([10], [10])):
        for my_default in (0.0, math.nan):
            for his_default in (0.0, math.nan):
                his_to_mine = []
                for mine, his in zip(my_len, his_len):
                    pool = list(range(mine)) + [None] * max(0, his - mine + 1)
                    picked = rng.permutation(len(pool))[:his]
                    his_to_mine.append([pool[i] for i in picked])
                yield dict(op="load", my_len=my_len, his_len=his_len, his_to_mine=his_to_mine,
                           my_default=my_default, his_default=his_default,
                           my_data=make_data(rng, int(np.prod(my_len)), my_default, 0.8, "int"),
                           his_data=make_data(rng, int(np.prod(his_len)), his_default, 0.6, "int"))


def load_linear_cases(seed=8):
    """The other store's innermost axis lands as one contiguous run in mine (his item j -> my
    item m0 + j): the vectorised scatter (128-bit when everything is 4-aligned, else rows)."""
    rng = np.random.default_rng(seed)
    for my_len, his_len, m0 in (([6, 12], [3, 8], 4), ([5, 7, 10], [4, 7, 6], 2), ([8, 16], [8, 16], 0),
                                ([9, 40, 24], [12, 33, 20], 4), ([50], [30], 8)):
        for my_default in (0.0, math.nan):
            for his_default in (0.0, math.nan):
                his_to_mine = []
                for d, (mine, his) in enumerate(zip(my_len, his_len)):
                    if d == len(my_len) - 1:
                        his_to_mine.append([m0 + j for j in range(his)])
                    else:
                        pool = list(range(mine)) + [None] * max(0, his - mine + 1)
                        picked = rng.permutation(len(pool))[:his]
                        his_to_mine.append([pool[i] for i in picked])
                yield dict(op="load", my_len=my_len, his_len=his_len, his_to_mine=his_to_mine,
                           my_default=my_default, his_default=his_default,
                           my_data=make_data(rng, int(np.prod(my_len)), my_default, 0.8, "int"),
                           his_data=make_data(rng, int(np.prod(his_len)), his_default, 0.6, "int"))


def run_case(case, store_cls):
    """Execute one case on a store class exposing the lowered interface; returns the
    resulting cells as float64 numpy (unset cells hold the default)."""
    op = case["op"]
    if op == "load":
        dst = store_cls(len(case["my_data"]), "float32", case["my_default"])
        _set(dst, case["my_data"])
        src = store_cls(len(case["his_data"]), "float32", case["his_default"])
        _set(src, case["his_data"])
        dst.load_lowered(src, case["my_len"], case["his_len"], case["his_to_mine"])
        return _get(dst)
    store = store_cls(len(case["data"]), case["type"], case["default"])
    _set(store, case["data"])
    if op == "drillUp":
        out = _call(store, "drillUp_lowered", case["old_len"], case["new_len"], case["maps"], case["method"])
    elif op == "drillDown":
        out = _call(store, "drillDown_lowered", case["old_len"], case["new_len"], case["maps"], case["method"],
                    case["distributions"])
    elif op == "dice":
        out = _call(store, "dice_lowered", case["old_len"], case["keep"])
    elif op == "reorder":
        out = _call(store, "reorder_lowered", case["old_len"], case["new_to_old"])
    else:
        raise AssertionError(op)
    return _get(out)


def _call(store, name, *args):
    fn = getattr(type(store), name)
    import inspect

    first = list(inspect.signature(fn).parameters)[0]
    if first == "stores":  # batched static form of the device store
        if name in ("drillUp_lowered",):
            return fn([store], args[0], args[1], args[2], [args[3]])[0]
        if name == "drillDown_lowered":
            return fn([store], args[0], args[1], args[2], [args[3]], [args[4]])[0]
        return fn([store], *args)[0]
    return fn(store, *args)


def _set(store, data):
    if hasattr(store, "set_data_f32"):
        store.set_data_f32(np.asarray(data, dtype=np.float32))
    else:
        store.data = [float(x) for x in data]


def _get(store):
    if hasattr(store, "data_f32"):
        return store.data_f32().astype(np.float64)
    if hasattr(store, "data_f64"):
        return store.data_f64()
    return np.asarray(store.data, dtype=np.float64)


def bits_equal(a, b):
    """float32 bit equality, NaN payload-insensitive."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(both_nan | (a.view(np.uint32) == b.view(np.uint32))))


def describe(case):
    keys = ("op", "old_len", "new_len", "method", "default", "type", "new_to_old", "my_len", "his_len")
    return ", ".join(f"{k}={case[k]}" for k in keys if k in case)
