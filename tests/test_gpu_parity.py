"""CUDA store vs the CPU oracle on seeded lowered cases (tests/cases.py), through the
C ABI.  Bars: bit-exact float32 for dice, reorder, load, presence,
first, last, highest, lowest; sum / average / drillDown are computed in double in the
reference's order, so they are expected bit-equal to fround(oracle) too (asserted), with
rel 1e-6 kept as the documented tolerance for tree-reduced regimes."""
import itertools
import math

import numpy as np
import pytest

import cases
from oracle.c_oracle import COracleStore
from oracle.store_oracle import OracleStore

pytestmark = pytest.mark.gpu

ALL = list(itertools.chain(cases.drillup_cases(), cases.drillup_long_cases(), cases.drilldown_cases(), cases.dice_cases(),
                           cases.reorder_cases(), cases.load_cases(), cases.load_linear_cases()))


def _gpu():
    from olap_in_memory_b200.store import GpuStore

    return GpuStore


@pytest.mark.parametrize("with_status", [True, False], ids=["status", "nostatus"])
@pytest.mark.parametrize("case", ALL, ids=[f"{i}-{c['op']}" for i, c in enumerate(ALL)])
def test_case_matches_oracle(case, with_status):
    G = _gpu()
    old = G.WITH_STATUS
    G.WITH_STATUS = with_status
    try:
        got = cases.run_case(case, G)
    finally:
        G.WITH_STATUS = old
    want = cases.run_case(case, COracleStore)
    assert cases.bits_equal(got, want.astype(np.float32)), (
        cases.describe(case) + f"\n got  {got[:16]}\n want {want[:16]}")


def test_appendix_a_edge_vectors():
    """SURVEY.md Appendix A1-A12, hand-derived from in-memory.js."""
    G = _gpu()
    nan = math.nan

    def up(data, default, method):
        s = G(len(data), "float32", default)
        s.set_data_f32(np.asarray(data, np.float32))
        return G.drillUp_lowered([s], [len(data)], [1], [np.zeros(len(data), np.int32)], [method])[0].data_f32()

    assert up([-1, 0, -2], 0, "highest").tolist() == [-1.0]           # A1
    assert up([2, 0, 4], 0, "average").tolist() == [3.0]              # A2
    assert up([0, 5, 7], 0, "first").tolist() == [5.0]                # A3
    assert up([2, 0, 3], 0, "product").tolist() == [6.0]              # A4
    assert up([1, -1, 5], 0, "average").tolist() == [np.float32(5 / 3)]  # A5
    assert up([10, 0, 20], nan, "average").tolist() == [10.0]         # A6
    assert np.isnan(up([1, nan], 0, "highest")[0]) and np.isnan(up([1, nan], 0, "lowest")[0])  # A7
    hi, lo = up([-0.0, 0.0], nan, "highest"), up([0.0, -0.0], nan, "lowest")  # A8
    assert hi[0] == 0 and not np.signbit(hi[0]) and lo[0] == 0 and np.signbit(lo[0])

    def down(data, default, typ, times):
        s = G(len(data), typ, default)
        s.set_data_f32(np.asarray(data, np.float32))
        m = np.repeat(np.arange(len(data), dtype=np.int32), times)
        return G.drillDown_lowered([s], [len(data)], [len(data) * times], [m], ["sum"])[0].data_f32()

    a9 = down([0, 6], nan, "float32", 2)
    assert np.isnan(a9[:2]).all() and a9[2:].tolist() == [3.0, 3.0]  # A9
    assert down([nan, 6], 0, "float32", 2).tolist() == [0.0, 0.0, 3.0, 3.0]  # A10
    assert down([-7], 0, "int32", 3).tolist() == [-3.0, -2.0, -3.0]  # A11
    s = G(2, "float32", 0)
    s.set_data_f32(np.asarray([10, 20], np.float32))
    out = G.drillDown_lowered([s], [2, 1], [2, 3], [np.arange(2, dtype=np.int32), np.zeros(3, np.int32)], ["sum"],
                              [[0.5, 0.3, 0.2, 0.1, 0.1, 0.8]])[0]
    assert np.allclose(out.data_f32(), [5, 3, 2, 2, 2, 16], rtol=1e-6)  # A12
    with pytest.raises(ValueError, match="distribution missing for index 4"):
        G.drillDown_lowered([s], [2, 1], [2, 3], [np.arange(2, dtype=np.int32), np.zeros(3, np.int32)], ["sum"],
                            [[0.5, 0.3, 0.2, 0.1, None, 0.8]])


def test_declared_divergences():
    """the device contract is ascending child index and 32-bit counts."""
    G = _gpu()
    s = G(3, "float32", 0)
    s.set_data_f32(np.asarray([10, 20, 30], np.float32))
    diced = G.dice_lowered([s], [3], [[2, 0, 1]])[0]
    assert diced.data_f32().tolist() == [30.0, 10.0, 20.0]
    zeros = [np.zeros(3, np.int32)]
    assert G.drillUp_lowered([diced], [3], [1], zeros, ["first"])[0].data_f32().tolist() == [30.0]  # reference: 10
    assert G.drillUp_lowered([diced], [3], [1], zeros, ["last"])[0].data_f32().tolist() == [20.0]   # reference: 30
    big = G(65536, "float32", 0)
    big.set_data_f32(np.full(65536, 2.0, np.float32))
    avg = G.drillUp_lowered([big], [65536], [1], [np.zeros(65536, np.int32)], ["average"])[0]
    assert avg.data_f32().tolist() == [2.0]  # reference: 131072 (Uint16 wrap)


def test_store_accessors_and_sparse_roundtrip():
    G = _gpu()
    rng = np.random.default_rng(5)
    for default in (0.0, math.nan):
        data = cases.make_data(rng, 10007, default, 0.3, "int")
        s = G(data.size, "float32", default)
        s.set_data_f32(data)
        o = OracleStore(data.size, "float32", default)
        o.data = [float(x) for x in data]
        assert cases.bits_equal(s.data_f32(), np.asarray(o.data, np.float32))
        assert s.presence().tolist() == [1 if i in o._dataMap else 0 for i in range(data.size)]
        assert s.status == o.status
        keys, vals = s.export_sparse()
        assert keys.tolist() == sorted(o._dataMap.keys())
        assert cases.bits_equal(vals, np.asarray([o._dataMap[k] for k in keys.tolist()], np.float32))
        if default == 0.0:
            finite = np.where(np.isnan(data), 0, data)
            assert not np.isnan(data).any() or np.isnan(s.total)
            if not np.isnan(data).any():
                assert math.isclose(s.total, float(finite.astype(np.float64).sum()), rel_tol=1e-9)
        else:
            assert math.isclose(s.total, o.total, rel_tol=1e-9)
        assert s.getValue(17) == np.float32(o.getValue(17)) or (math.isnan(s.getValue(17)) and math.isnan(o.getValue(17)))
        s.setValue(17, 123.5)
        o.setValue(17, 123.5)
        s.setValue(18, None)
        o.setValue(18, None)
        assert s.getValue(17) == 123.5 and cases.bits_equal(s.data_f32(), np.asarray(o.data, np.float32))
        c = s.clone()
        s.fill(7.0)
        assert c.getValue(17) == 123.5 and s.getValue(17) == 7.0 and s.total == 7.0 * data.size
        r = G(data.size, "float32", default)
        r.import_sparse(*c.export_sparse())
        assert cases.bits_equal(r.data_f32(), c.data_f32())
    with pytest.raises(ValueError, match=r"value length is invalid: 4 !== 3"):
        G(4).set_data_f32(np.zeros(3, np.float32))
    with pytest.raises(ValueError, match="Invalid default value"):
        G(4, "float32", 1)
    with pytest.raises(ValueError, match="Invalid type"):
        G(4, "float16", 0)
    assert G(6, "uint32", 0).byteLength == 24 and G(6, "float64", 0).byteLength == 48
    assert G(0).data == []


def test_batched_measures_share_one_call():
    """All stored measures of a cube in ONE olap_drill_up call, mixed methods."""
    G = _gpu()
    rng = np.random.default_rng(9)
    dims, P = [40, 24], 6
    m = cases.random_map(rng, 40, P, True)
    methods = ["sum", "average", "highest", "lowest", "first", "last", "product"]
    datas = [cases.make_data(rng, 960, 0.0, 0.7, "small" if k == "product" else "int") for k in methods]
    stores = []
    for d in datas:
        s = G(960, "float32", 0.0)
        s.set_data_f32(d)
        stores.append(s)
    outs = G.drillUp_lowered(stores, dims, [P, 24], [m, cases.identity(24)], methods)
    for d, method, out in zip(datas, methods, outs):
        o = COracleStore(960, "float32", 0.0)
        o.set_data_f32(d)
        want = o.drillUp_lowered(dims, [P, 24], [m, cases.identity(24)], method).data_f64()
        assert cases.bits_equal(out.data_f32(), want.astype(np.float32)), method


def test_status_plane_semantics():
    """README.md:698-721: drillUp ORs children (0x3 = incomplete), drillDown adds 0x4."""
    G = _gpu()
    s = G(6, "float32", math.nan)
    s.set_data_f32(np.asarray([1, math.nan, 3, math.nan, math.nan, 4], np.float32))
    assert s.status == [2, 1, 2, 1, 1, 2]
    up = G.drillUp_lowered([s], [3, 2], [3, 1], [cases.identity(3), np.zeros(2, np.int32)], ["sum"])[0]
    assert up.status == [3, 3, 3] and up.data_f32().tolist() == [1.0, 3.0, 4.0]
    up2 = G.drillUp_lowered([s], [3, 2], [1, 2], [np.zeros(3, np.int32), cases.identity(2)], ["sum"])[0]
    assert up2.status == [3, 3]
    dn = G.drillDown_lowered([up], [3], [6], [np.repeat(np.arange(3, dtype=np.int32), 2)], ["sum"])[0]
    assert dn.status == [7, 7, 7, 7, 7, 7]


def test_computed_measures():
    G = _gpu()
    from olap_in_memory_b200 import Cube, GenericDimension

    rng = np.random.default_rng(11)
    formulas = ["a + b", "a - b * c", "(a + b) / c", "a || b", "isNaN(a) + 2 * b", "a / a__total + max(b, c, 3)",
                "min(a, b) ^ 2 % 7", "-a + abs(b - c)", "a ? b : c", "sqrt(abs(a)) + round(b / 3) + floor(c / 7)"]
    for default in (0.0, math.nan):
        def mk(store_cls):
            cube = Cube([GenericDimension("x", "root", [str(i) for i in range(37)]),
                         GenericDimension("y", "root", [str(i) for i in range(29)])], store_cls)
            r = np.random.default_rng(3)
            for name in ("a", "b", "c"):
                cube.createStoredMeasure(name * 2, {}, "float32", default)
                cube.setData(name * 2, cases.make_data(r, cube.storeSize, default, 0.8, "int").tolist())
            for k, f in enumerate(formulas):
                cube.createComputedMeasure(f"f{k}_m", re_sub(f))
            return cube

        def re_sub(f):
            import re

            return re.sub(r"\b([abc])\b", lambda mo: mo.group(1) * 2, f).replace("a__total", "aa__total")

        gpu, cpu = mk(G), mk(OracleStore)
        for k, f in enumerate(formulas):
            got = np.asarray(gpu.getData(f"f{k}_m"))
            want = np.asarray(cpu.getData(f"f{k}_m"))
            assert np.allclose(got, want, rtol=1e-6, atol=0, equal_nan=True), (f, default)


def test_kernel_paths_are_the_intended_ones():
    """Shape classes of SURVEY.md Appendix B reach the kernel written for them."""
    from olap_in_memory_b200 import _native as N

    G = _gpu()
    path = lambda: N.lib().olap_last_op_path().decode()
    s = G(64 * 40 * 64, "float32", 0)
    ident = cases.identity
    m = cases.random_map(np.random.default_rng(0), 40, 5, True)
    G.drillUp_lowered([s], [64, 40, 64], [64, 5, 64], [ident(64), m, ident(64)], ["sum"])
    assert path() == "drillup/mid-vec4"
    G.drillUp_lowered([s], [64 * 64, 40], [64 * 64, 5], [ident(64 * 64), m], ["sum"])
    assert path() == "drillup/tile"
    G.drillUp_lowered([s], [64, 40, 64], [2, 5, 64], [cases.random_map(np.random.default_rng(1), 64, 2, False), m, ident(64)], ["sum"])
    assert path() == "drillup/generic"
    G.reorder_lowered([s], [64, 40, 64], [1, 0, 2])
    assert path() == "gather/vec4"
    G.reorder_lowered([s], [64, 40, 64], [2, 1, 0])
    assert path() in ("reorder/pair-async", "reorder/pair-transpose")
    G.reorder_lowered([G(63 * 40 * 65, "float32", 0)], [63, 40, 65], [2, 1, 0])
    assert path() == "reorder/box-transpose"
    G.dice_lowered([s], [64, 40, 64], [ident(64), np.arange(0, 40, 2, dtype=np.int32), ident(64)])
    assert path() == "gather/vec4"


def test_inf_minus_inf_restart_under_nan_default():
    """in-memory.js:311-318: under a NaN default inf + -inf = NaN deletes the key and the next
    set child restarts the accumulator — on every drillUp kernel path."""
    G = _gpu()
    inf = math.inf
    column = [3.0, inf, -inf, 5.0, 7.0, math.nan, -inf, inf]  # -> restart at 5: 12; then -inf+inf deletes again
    for C_, I_, dims_of in ((8, 64, lambda: ([8, 64], 0)), (8, 1, lambda: ([5, 8], 1)), (8, 6, lambda: ([8, 6], 0))):
        old_len, d = dims_of()
        n = int(np.prod(old_len))
        data = np.full(n, 2.0, np.float32).reshape(old_len)
        column2 = [inf, -inf, 5.0, 7.0, 1.0, math.nan, 1.0, 1.0]  # restart at 5 -> 15
        if d == 0:
            data[:, 0] = column
            data[:, 1] = column2
        else:
            data[0, :] = column
            data[1, :] = column2
        for method in ("sum", "average"):
            s = G(n, "float32", math.nan)
            s.set_data_f32(data.ravel())
            o = COracleStore(n, "float32", math.nan)
            o.set_data_f32(data.ravel())
            new_len = list(old_len)
            new_len[d] = 1
            maps = [cases.identity(x) for x in old_len]
            maps[d] = np.zeros(old_len[d], np.int32)
            got = G.drillUp_lowered([s], old_len, new_len, maps, [method])[0].data_f32()
            want = o.drillUp_lowered(old_len, new_len, maps, method).data_f64()
            assert cases.bits_equal(got, want.astype(np.float32)), (old_len, method, got[:4], want[:4])


def test_more_measures_than_fit_in_kernel_parameters():
    """> 16 measures in one olap_drill_up call take the device-table path for descriptors."""
    G = _gpu()
    rng = np.random.default_rng(21)
    dims, P, n_meas = [30, 40], 4, 20
    m = cases.random_map(rng, 30, P, False)
    methods = [cases.METHODS[k % 6] for k in range(n_meas)]
    datas = [cases.make_data(rng, 1200, 0.0, 0.7, "int") for _ in range(n_meas)]
    stores = []
    for d in datas:
        s = G(1200, "float32", 0.0)
        s.set_data_f32(d)
        stores.append(s)
    outs = G.drillUp_lowered(stores, dims, [P, 40], [m, cases.identity(40)], methods)
    for d, method, out in zip(datas, methods, outs):
        o = COracleStore(1200, "float32", 0.0)
        o.set_data_f32(d)
        want = o.drillUp_lowered(dims, [P, 40], [m, cases.identity(40)], method).data_f64()
        assert cases.bits_equal(out.data_f32(), want.astype(np.float32)), method


def test_shared_status_plane_layout():
    """north-star layout (a): the measures of a cube in ONE allocation with ONE status plane
    shared by all of them (olap_store_create_batch(..., shared_status=1)); transforms keep the
    sharing and touch the plane once."""
    import ctypes as C

    from olap_in_memory_b200 import _native as N

    G = _gpu()
    lib = N.lib()
    n, size = 3, 6 * 8
    handles = (C.c_void_p * n)()
    types = (C.c_int * n)(2, 2, 2)
    kinds = (C.c_int * n)(0, 0, 0)
    N.check(lib.olap_store_create_batch(n, size, types, kinds, 1, 1, handles))
    stores = [G._wrap(handles[k]) for k in range(n)]
    ptrs = {lib.olap_store_status_ptr(s._h) for s in stores}
    assert len(ptrs) == 1 and None not in ptrs
    v0 = lib.olap_store_values_ptr(stores[0]._h)
    assert lib.olap_store_values_ptr(stores[1]._h) - v0 == 256  # 48 floats = 192 B padded to 256
    data = np.arange(1, size + 1, dtype=np.float32)
    data[5] = 0.0  # same fill pattern for every measure: that is when one plane is exact
    for k, s in enumerate(stores):
        s.set_data_f32(data * (k + 1))
    assert stores[2].status[5] == 1 and stores[0].status[4] == 2
    m = np.asarray([0, 0, 1, 1, 2, 2], np.int32)
    outs = G.drillUp_lowered(stores, [6, 8], [3, 8], [m, cases.identity(8)], ["sum", "highest", "last"])
    assert len({lib.olap_store_status_ptr(o._h) for o in outs}) == 1
    x = data.reshape(6, 8)
    assert outs[0].data_f32().reshape(3, 8).tolist() == (x[0::2] + x[1::2]).tolist()
    assert outs[1].data_f32().reshape(3, 8).tolist() == (2 * np.maximum(x[0::2], x[1::2])).tolist()
    assert outs[2].data_f32().reshape(3, 8).tolist() == (3 * x[1::2]).tolist()
    st = np.asarray(outs[1].status).reshape(3, 8)
    assert st[0, 5] == 3 and (np.delete(st.ravel(), 5) == 2).all()


def test_cluster_transpose(monkeypatch):
    """OLAP_PAIR_CLUSTER=1: the 2-CTA cluster / distributed-shared-memory variant of the pair
    transpose (off by default, see kernels_pair.cuh) stays bit-exact, ragged tiles included."""
    G = _gpu()
    monkeypatch.setenv("OLAP_PAIR_CLUSTER", "1")
    rng = np.random.default_rng(5)
    for dims, perm in (([24, 16, 10, 10, 10], [4, 3, 2, 1, 0]), ([28, 12, 6, 10, 10], [4, 3, 2, 1, 0]),
                       ([400, 408], [1, 0]), ([20, 20, 20, 10, 10, 10], [5, 4, 3, 2, 1, 0])):
        n = int(np.prod(dims))
        data = rng.integers(1, 1000, n).astype(np.float32)
        data[rng.random(n) < 0.3] = 0.0
        s = G(n, "float32", 0)
        s.set_data_f32(data)
        out = G.reorder_lowered([s], dims, perm)[0]
        want = data.reshape(dims).transpose(perm).reshape(-1)
        assert np.array_equal(out.data_f32(), want)
        st = out.status
        if st is not None:
            assert np.array_equal(np.asarray(st), np.where(want != 0, 2, 1))


def test_reorder_fuzz_against_numpy():
    """reorder is a pure permutation: random shapes / permutations (4-aligned and not, so that
    the pair transpose, the box transpose and the vector gather all get hit, ragged tiles
    included), two measures per call, against numpy.transpose — values and status bytes."""
    from olap_in_memory_b200 import _native as N

    G = _gpu()
    rng = np.random.default_rng(11)
    paths = {}
    for trial in range(70):
        nd = int(rng.integers(2, 6))
        if trial % 2:
            dims = [int(rng.choice([4, 8, 12, 20, 36, 44, 100, 104, 200])) for _ in range(nd)]
        else:
            dims = [int(rng.integers(2, 60)) for _ in range(nd)]
        while int(np.prod(dims)) > 3_000_000:
            dims[int(np.argmax(dims))] //= 2
        perm = [int(x) for x in rng.permutation(nd)]
        n = int(np.prod(dims))
        stores, datas = [], []
        for _ in range(2):
            data = rng.integers(1, 1000, n).astype(np.float32)
            data[rng.random(n) < 0.25] = 0.0
            s = G(n, "float32", 0)
            s.set_data_f32(data)
            stores.append(s)
            datas.append(data)
        # measure 0: status plane derived from the values (never read); measure 1: a mutable raw pointer was handed
        # out, so its plane is loaded and moved — both flavours of every kernel in one call
        from olap_in_memory_b200 import interop
        interop.status_tensor(stores[1])
        assert stores[0].status_derived and not stores[1].status_derived
        outs = G.reorder_lowered(stores, dims, perm)
        assert outs[0].status_derived and not outs[1].status_derived
        path = N.lib().olap_last_op_path().decode()
        paths[path] = paths.get(path, 0) + 1
        for data, out in zip(datas, outs):
            want = data.reshape(dims).transpose(perm).reshape(-1)
            assert np.array_equal(out.data_f32(), want), (dims, perm, path)
            assert np.array_equal(np.asarray(out.status), np.where(want != 0, 2, 1)), (dims, perm, path)
    disjoint = paths.get("reorder/pair-transpose", 0) + paths.get("reorder/pair-async", 0) + paths.get("reorder/tma-transpose", 0)
    assert disjoint >= 5 and paths.get("reorder/box-transpose", 0) >= 5, paths


@pytest.mark.parametrize("knobs", [{}, {"OLAP_TMA_OUT": "200"}, {"OLAP_TMA_STAGES": "1", "OLAP_TMA_STORE_STAGES": "1"}])
def test_tma_transpose(monkeypatch, knobs):
    """The tensor-map variant of the disjoint-group transpose (kernels_tma.cuh: cp.async.bulk.tensor
    loads into a ring of shared-memory stages, register transposition, tensor-map stores; status
    bytes on the side): bit-exact against numpy.transpose, ragged tiles (TMA zero-fill / clipping),
    merged dimensions, several measures, with and without status plane, a NaN default."""
    from olap_in_memory_b200 import _native as N

    G = _gpu()
    monkeypatch.setenv("OLAP_TRANSPOSE_TMA", "1")  # opt-in: measured slower than the pair kernel (profiles/README.md)
    for key, value in knobs.items():
        monkeypatch.setenv(key, value)
    rng = np.random.default_rng(23)
    hit = 0
    for dims, perm in (([24, 16, 10, 10, 10], [4, 3, 2, 1, 0]), ([28, 12, 6, 10, 10], [4, 3, 2, 1, 0]),
                       ([400, 408], [1, 0]), ([20, 20, 20, 10, 10, 10], [5, 4, 3, 2, 1, 0]),
                       ([36, 52, 44], [2, 1, 0]), ([104, 9, 100], [2, 1, 0]), ([12, 200, 3, 100], [3, 2, 0, 1]),
                       ([3652, 32, 32], [2, 1, 0]), ([100, 100, 12, 12], [2, 3, 0, 1])):
        n = int(np.prod(dims))
        for with_status, default in ((True, 0.0), (False, 0.0), (True, float("nan"))):
            stores, datas = [], []
            for _ in range(2):
                data = rng.integers(1, 1000, n).astype(np.float32)
                data[rng.random(n) < 0.3] = default
                s = G(n, "float32", default, with_status=with_status)
                s.set_data_f32(data)
                stores.append(s)
                datas.append(data)
            outs = G.reorder_lowered(stores, dims, perm)
            hit += N.lib().olap_last_op_path() == b"reorder/tma-transpose"
            for data, out in zip(datas, outs):
                want = data.reshape(dims).transpose(perm).reshape(-1)
                assert np.array_equal(out.data_f32().view(np.uint32), want.view(np.uint32)), (dims, perm, with_status, default)
                if with_status:
                    set_ = want == want if default != default else want != 0
                    assert np.array_equal(np.asarray(out.status), np.where(set_, 2, 1)), (dims, perm)
    assert hit >= 12, hit


def test_integer_stores_refuse_values_their_float32_cell_would_change():
    """the reference keeps JS doubles for every store type; an int32 / uint32 store whose
    count exceeds 2^24 must not silently come back as a neighbouring integer."""
    from olap_in_memory_b200 import _native as N

    G = _gpu()
    s = G(4, "uint32", 0)
    s.data = [1.0, 16777216.0, 3.0, 0.5]  # exact in Float32 (non-integers are the caller's business, as in JS)
    assert s.data == [1.0, 16777216.0, 3.0, 0.5]
    with pytest.raises(N.OlapError, match="16777217 at index 2 of an uint32 store"):
        s.data = [1.0, 2.0, 16777217.0, 4.0]
    with pytest.raises(N.OlapError, match="not representable"):
        s.setValue(1, 16777217.0)
    with pytest.raises(N.OlapError, match="int32 store"):
        G(2, "int32", 0).setValues([0, 1], [5.0, -33554433.0])
    f = G(2, "float64", 0)  # float64 stores round by declared contract
    f.data = [0.1, 16777217.0]
    assert f.data == [float(np.float32(0.1)), 16777216.0]


def test_derived_status_planes_are_tracked_and_not_needed():
    """A store whose status plane follows from its values (after set data / fill / sparse import / eval, through
    dice, reorder, clone) is rolled up WITHOUT reading that plane (4 instead of 5 bytes per input cell); the
    flag must be lost exactly where the plane starts to carry information of its own."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200 import interop

    G = _gpu()
    lib = N.lib()
    rng = np.random.default_rng(3)
    C_, I = 12, 64
    data = rng.integers(1, 100, C_ * I).astype(np.float32)
    data[rng.random(C_ * I) < 0.4] = 0.0
    s = G(C_ * I, "float32", 0)
    assert s.status_derived  # freshly created: every cell UNSET and default
    s.set_data_f32(data)
    assert s.status_derived
    gmap = (np.arange(C_) // 4).astype(np.int32)
    up = G.drillUp_lowered([s], [C_, I], [3, I], [gmap, None], ["sum"])[0]
    assert lib.olap_last_op_path() == b"drillup/mid-vec4" and not up.status_derived  # OR-merged flags (0x3 = incomplete)
    # the same rollup with the plane READ (a mutable raw pointer was handed out: the library forgets)
    t = G(C_ * I, "float32", 0)
    t.set_data_f32(data)
    interop.status_tensor(t)
    assert not t.status_derived
    up2 = G.drillUp_lowered([t], [C_, I], [3, I], [gmap, None], ["sum"])[0]
    assert np.array_equal(up.data_f32(), up2.data_f32()) and up.status == up2.status
    assert set(up.status) <= {1, 2, 3} and 3 in up.status
    t.canonicalise()
    assert t.status_derived
    # through dice / reorder / clone the flag survives; through drillDown it does not
    d = G.dice_lowered([s], [C_, I], [np.arange(0, C_, 2, dtype=np.int32), np.arange(I, dtype=np.int32)])[0]
    r = G.reorder_lowered([s], [C_, I], [1, 0])[0]
    assert d.status_derived and r.status_derived and s.clone().status_derived
    down = G.drillDown_lowered([up], [3, I], [C_, I], [gmap, np.arange(I, dtype=np.int32)], ["sum"])[0]
    assert not down.status_derived
    # an interpolated cube rolled up again keeps its 0x4 flags: the plane IS read there
    again = G.drillUp_lowered([down], [C_, I], [3, I], [gmap, None], ["sum"])[0]
    assert any(b & 4 for b in again.status)


def test_readme_only_status_examples():
    """README-ONLY (parity unpinned: this fork of the reference has no getStatus).  The examples of
    /root/reference/README.md:698-732 and :755-771 transcribed as data: flags 0x1 not set / 0x2 set / 0x4
    interpolated; `is_empty` = (s & 3) == 1, `is_incomplete` = (s & 3) == 3, `is_complete` = (s & 3) == 2,
    `is_interpolated` = (s & 4) == 4; c === complete === !(status & 0x1), r === raw === !(status & 0x4)."""
    from olap_in_memory_b200 import Cube, GenericDimension, TimeDimension

    _gpu()

    def flags(status):
        return {"is_empty": (status & 0x3) == 0x1, "is_incomplete": (status & 0x3) == 0x3, "is_complete": (status & 0x3) == 0x2,
                "is_interpolated": (status & 0x4) == 0x4}

    # README.md:722-727: [NaN, empty], [1, incomplete], [2, complete] — a month cube rolled up to quarters:
    # Q1 has no month set, Q2 has one of three, Q3 all three
    cube = Cube([TimeDimension("time", "month", "2010-01", "2010-09")])
    cube.createStoredMeasure("main_measure", {"time": "sum"}, "float32", math.nan)
    nan = math.nan
    cube.setData("main_measure", [nan, nan, nan, 1, nan, nan, 0.5, 0.5, 1])
    q = cube.drillUp("time", "quarter")
    data, status = q.getData("main_measure"), q.getStatus("main_measure")
    assert math.isnan(data[0]) and data[1:] == [1.0, 2.0]
    assert [flags(s) for s in status] == [
        {"is_empty": True, "is_incomplete": False, "is_complete": False, "is_interpolated": False},
        {"is_empty": False, "is_incomplete": True, "is_complete": False, "is_interpolated": False},
        {"is_empty": False, "is_incomplete": False, "is_complete": True, "is_interpolated": False},
    ]
    # README.md:755-771: every cell set and raw -> { v, r: true, c: true }; after a drillDown r turns false
    cube = Cube([GenericDimension("city", "root", ["paris", "madrid"]), TimeDimension("time", "quarter", "2010-Q1", "2010-Q2")])
    cube.createStoredMeasure("main_measure", {"time": "sum"}, "float32", math.nan)
    cube.setData("main_measure", [33, 7, 33, 7])
    st = cube.getStatus("main_measure")
    assert [(v, not (s & 0x4), not (s & 0x1)) for v, s in zip(cube.getData("main_measure"), st)] == [(33.0, True, True), (7.0, True, True)] * 2
    months = cube.drillDown("time", "month")
    assert all((s & 0x4) and not (s & 0x1) for s in months.getStatus("main_measure"))  # interpolated, still complete
    assert months.getData("main_measure")[:3] == [11.0, 11.0, 11.0]
    # "When the cube is filled with .hydrateFromCube(otherCube), the status flags are copied between cubes" (README.md:704)
    target = Cube([GenericDimension("city", "root", ["paris", "madrid"]), TimeDimension("time", "month", "2010-01", "2010-06")])
    target.createStoredMeasure("main_measure", {"time": "sum"}, "float32", math.nan)
    target.hydrateFromCube(months)
    assert all(s & 0x4 for s in target.getStatus("main_measure"))


@pytest.mark.parametrize("geometry", ["0", "1"], ids=["tile64x2", "tile128x1"])
@pytest.mark.parametrize("default", [0.0, math.nan])
@pytest.mark.parametrize("monotone", [False, True])
def test_lanes_rollup_of_many_long_rows(default, monotone, geometry, monkeypatch):
    """drillup_lanes_kernel ([O >= 64, C, 1] -> [O, P <= 8, 1], rows too long for a tile, any map: customers -> segment):
    the row sits on the lane, the warp-uniform parent selects the accumulator.  Against the C oracle: order-only rules
    bit-exact (first / last follow child order through the ordered folds), sums within 1e-6 (tree-reduced regime),
    status bytes exact, with a loaded and with a derived status plane, one and several segments per row, ragged row
    groups and tiles, a parent without children.  A loaded plane whose rows do not start on 4 bytes stays with the long kernel."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200 import interop
    from oracle.c_oracle import COracleStore

    G = _gpu()
    # both pipeline geometries, and the loaded-plane path (by default such sources stay with the long kernel)
    monkeypatch.setenv("OLAP_LANES_GEO", geometry)
    monkeypatch.setenv("OLAP_LANES_LOADED", "1")
    rng = np.random.default_rng(9)
    # rows of more than 200 KB: shorter ones fit a shared-memory tile and belong to the tile kernel
    for O, C_, P, ss in ((70, 52304, 5, None), (129, 52048 + 76, 8, None), (64, 59000, 1, "1"), (97, 52600, 3, "1"),
                         (65, 52052, 2, "2"), (70, 52301, 5, None)):
        if ss is None:
            monkeypatch.delenv("OLAP_LANES_SS", raising=False)
        else:
            monkeypatch.setenv("OLAP_LANES_SS", ss)
        cmap = cases.random_map(rng, C_, P, monotone)
        if P == 5:
            cmap[cmap == 3] = 2  # parent 3 has no child at all: its cells stay unset
        ident = np.arange(O, dtype=np.int32)
        for kind in ("int", "pow2"):
            if kind == "pow2":
                # products of 0.5 / 1 / 2 stay far from underflow over 10^4 children (a product that lands on the
                # default restarts in the reference, which no split reduction can follow) and are exact in any order
                data = np.exp2(rng.integers(-1, 2, O * C_)).astype(np.float32)
                data[rng.random(O * C_) >= 0.6] = default
            else:
                data = cases.make_data(rng, O * C_, default, 0.6, kind)
            methods = ["sum", "average", "highest", "lowest", "first", "last"] + (["product"] if kind == "pow2" else [])
            ref = COracleStore(O * C_, "float32", default)
            ref.set_data_f32(data)
            for derived in (True, False):
                stores = []
                for _ in methods:
                    s = G(O * C_, "float32", default)
                    s.set_data_f32(data)
                    if not derived:
                        interop.status_tensor(s)
                    stores.append(s)
                outs = G.drillUp_lowered(stores, [O, C_], [O, P], [ident, cmap], methods)
                # a plane that has to be read travels in words of 4 status bytes: rows must start on them
                assert N.lib().olap_last_op_path() == (b"drillup/lanes" if derived or C_ % 4 == 0 else b"drillup/long"), (O, C_, P)
                for method, out in zip(methods, outs):
                    want = ref.drillUp_lowered([O, C_], [O, P], [ident, cmap], method).data_f64().astype(np.float32)
                    got = out.data_f32()
                    if method in ("sum", "average", "product"):
                        assert np.allclose(got, want, rtol=1e-6, atol=0, equal_nan=True), (O, C_, P, kind, method, derived)
                    else:
                        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (O, C_, P, kind, method, derived)
                # status: OR of the children's bytes, UNSET for a parent without children
                set_ = (data == data) if default != default else (data != 0)
                bits = np.where(set_, 2, 1).astype(np.uint8).reshape(O, C_)
                want_st = np.ones((O, P), dtype=np.uint8)
                for q in range(P):
                    cols = np.flatnonzero(cmap == q)
                    if cols.size:
                        want_st[:, q] = np.bitwise_or.reduce(bits[:, cols], axis=1)
                assert np.array_equal(np.asarray(outs[0].status, dtype=np.uint8).reshape(O, P), want_st), (O, C_, P, kind, derived)


def test_pageable_host_buffers_travel_through_the_pinned_ring():
    """Large transfers from / to plain (pageable) host memory are cut into chunks that go through a ring of pinned
    buffers, the host-side copies on worker threads (csrc/host_pipe.cuh; in-memory.js:30-46 `get data` / `set data`).
    Sizes that wrap the ring several times and end on a ragged chunk: every byte must arrive, in both directions, for
    Float32 and Float64 arrays, status and presence bytes and the sparse lists."""
    from olap_in_memory_b200 import _native as N

    G = _gpu()
    rng = np.random.default_rng(21)
    n = 5 * (16 << 20) // 4 + 12345  # 5 chunks of 16 MiB and a tail
    data = rng.integers(-5, 6, n).astype(np.float32)
    data[rng.random(n) < 0.3] = 0.0
    s = G(n, "float32", 0)
    s.set_data_f32(data)
    assert np.array_equal(s.data_f32().view(np.uint32), data.view(np.uint32))
    st = np.empty(n, dtype=np.uint8)
    N.check(N.lib().olap_store_status(s._h, st.ctypes.data, n))
    assert np.array_equal(st, np.where(data != 0, 2, 1).astype(np.uint8))
    assert np.array_equal(s.presence(), (data != 0).astype(np.uint8))
    keys, values = s.export_sparse()
    want_keys = np.flatnonzero(data != 0)
    assert np.array_equal(keys, want_keys) and np.array_equal(values, data[want_keys])
    t = G(n, "float32", 0)
    t.import_sparse(keys, values)
    assert np.array_equal(t.data_f32().view(np.uint32), data.view(np.uint32))
    # Float64 in and out (upload_f64 / download_f64)
    d64 = data.astype(np.float64)
    u = G(n, "float64", math.nan)
    N.check(N.lib().olap_store_upload_f64(u._h, d64.ctypes.data, n))
    out64 = np.empty(n, dtype=np.float64)
    N.check(N.lib().olap_store_download_f64(u._h, out64.ctypes.data, n))
    assert np.array_equal(out64, d64)
    assert s.total == float(d64.sum())


@pytest.mark.parametrize("default", [0.0, math.nan])
def test_dice_of_the_innermost_axis_alone(default):
    """gather_inner_flat_kernel: [R, D] -> [R, K] with every other axis untouched (dice / slice of the innermost
    dimension, in-memory.js:213-263): staged spans of whole rows, kept cells picked out of shared memory.  Bit-exact
    against numpy fancy indexing for every-other, reordered, single-item and ragged lists, row counts that do not fill
    the last CTA, outer axes that merge into the row count, loaded and derived status planes."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200 import interop

    G = _gpu()
    rng = np.random.default_rng(33)
    shapes = [([1000, 10], list(range(0, 10, 2))), ([77, 13, 10], [9, 0, 4]), ([4099, 7], [3]), ([300, 33], list(range(32, -1, -1))),
              ([20, 50, 512], sorted(rng.choice(512, 100, replace=False).tolist())), ([6000, 4], [1, 2])]
    for lens, keep in shapes:
        n = int(np.prod(lens))
        data = cases.make_data(rng, n, default, 0.6, "int")
        keeps = [np.arange(d, dtype=np.int32) for d in lens[:-1]] + [np.asarray(keep, np.int32)]
        want = data.reshape(-1, lens[-1])[:, keep].reshape(-1)
        set_ = (data == data) if default != default else (data != 0)
        want_st = np.where(set_, 2, 1).astype(np.uint8).reshape(-1, lens[-1])[:, keep].reshape(-1)
        for derived in (True, False):
            s = G(n, "float32", default)
            s.set_data_f32(data)
            if not derived:
                st = interop.status_tensor(s)  # a plane of its own content: the kernel has to move it
                st[::3] |= 4
                want_st_l = np.where(set_, 2, 1).astype(np.uint8)
                want_st_l[::3] |= 4
                want_st_l = want_st_l.reshape(-1, lens[-1])[:, keep].reshape(-1)
            out = G.dice_lowered([s], lens, keeps)[0]
            if len(keep) > 1:  # a single kept item is a plain strided gather
                assert N.lib().olap_last_op_path() == b"gather/inner-flat", (lens, keep)
            assert np.array_equal(out.data_f32().view(np.uint32), want.view(np.uint32)), (lens, keep, derived)
            assert np.array_equal(np.asarray(out.status, np.uint8), want_st if derived else want_st_l), (lens, keep, derived)
            assert bool(out.status_derived) == derived


@pytest.mark.parametrize("default", [0.0, math.nan])
def test_rearrangements_inside_short_blocks(default):
    """gather_inner_flat_kernel beyond the single diced axis: reorders that swap trailing axes (the 10 x 10 inner swap
    of config 3, a 3-axis rotation of a 6 x 5 x 4 block) and dices of two trailing axes at once, leading axes
    untouched — against numpy transpose / fancy indexing, values and status bytes, loaded and derived planes."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200 import interop

    G = _gpu()
    rng = np.random.default_rng(35)
    # inner-flat: the block stays innermost; flat-to-front: a short innermost axis (or block) becomes the OUTERMOST part
    # of the output (one run per plane leaves a tile); planes-to-inner: the mirror image, short outer axes become the
    # innermost ones
    flat, front, planes = b"gather/inner-flat", b"gather/flat-to-front", b"gather/planes-to-inner"
    for kind, lens, arg, want_path in (
            ("reorder", [70, 9, 10, 10], [0, 1, 3, 2], flat), ("reorder", [1000, 6, 5, 4], [0, 3, 1, 2], flat),
            ("reorder", [333, 7, 3], [0, 2, 1], flat), ("dice", [500, 10, 12], [[1, 3, 8], [0, 5, 11, 2]], flat),
            ("dice", [90, 4, 10, 10], [None, [9, 0], list(range(10))], flat),
            ("reorder", [300, 7, 10], [2, 0, 1], front), ("reorder", [2001, 4, 3], [1, 2, 0], front), ("reorder", [40, 50, 3, 2], [3, 2, 0, 1], front),
            ("reorder", [10, 5004], [1, 0], planes), ("reorder", [3, 4, 5000], [2, 0, 1], planes), ("reorder", [2, 70, 100], [1, 2, 0], planes)):
        n = int(np.prod(lens))
        data = cases.make_data(rng, n, default, 0.6, "int")
        set_ = (data == data) if default != default else (data != 0)
        st_derived = np.where(set_, 2, 1).astype(np.uint8)
        st_loaded = st_derived.copy()
        st_loaded[::3] |= 4

        def move(x):
            x = x.reshape(lens)
            if kind == "reorder":
                return np.ascontiguousarray(x.transpose(arg)).reshape(-1)
            keeps = [np.arange(d) if k is None else np.asarray(k) for d, k in zip(lens, [None] * (len(lens) - len(arg)) + arg)]
            return np.ascontiguousarray(x[np.ix_(*keeps)]).reshape(-1)

        for derived in (True, False):
            s = G(n, "float32", default)
            s.set_data_f32(data)
            if not derived:
                interop.status_tensor(s)[::3] |= 4
            if kind == "reorder":
                out = G.reorder_lowered([s], lens, arg)[0]
            else:
                full = [None] * (len(lens) - len(arg)) + arg
                out = G.dice_lowered([s], lens, [np.arange(d, dtype=np.int32) if k is None else np.asarray(k, np.int32) for d, k in zip(lens, full)])[0]
            assert N.lib().olap_last_op_path() == want_path, (kind, lens, arg)
            assert np.array_equal(out.data_f32().view(np.uint32), move(data).view(np.uint32)), (kind, lens, arg, derived)
            assert np.array_equal(np.asarray(out.status, np.uint8), move(st_derived if derived else st_loaded)), (kind, lens, arg, derived)


def test_async_pair_transpose(monkeypatch):
    """OLAP_PAIR_ASYNC=1: the cp.async-staged variant of the disjoint-group transpose (kernels_pair_async.cuh;
    opt-in, measured slower than the register-staged kernel): bit-exact against numpy.transpose, ragged tiles,
    loaded and derived status planes in one call."""
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200 import interop

    G = _gpu()
    monkeypatch.setenv("OLAP_PAIR_ASYNC", "1")
    rng = np.random.default_rng(41)
    taken = 0
    for dims, perm in (([64, 40, 64], [2, 1, 0]), ([100, 12, 36, 10, 10], [4, 3, 2, 1, 0]), ([132, 3, 72], [2, 1, 0]), ([8, 200, 104], [0, 2, 1])):
        n = int(np.prod(dims))
        datas = [cases.make_data(rng, n, 0.0, 0.6, "int") for _ in range(2)]
        stores = []
        for data in datas:
            s = G(n, "float32", 0)
            s.set_data_f32(data)
            stores.append(s)
        interop.status_tensor(stores[1])[::5] |= 4  # a plane of its own content
        outs = G.reorder_lowered(stores, dims, perm)
        taken += N.lib().olap_last_op_path() == b"reorder/pair-async"  # shapes the pair planner declines take other kernels
        for k, (data, out) in enumerate(zip(datas, outs)):
            want = data.reshape(dims).transpose(perm).reshape(-1)
            assert np.array_equal(out.data_f32().view(np.uint32), want.view(np.uint32)), (dims, perm, k)
            st = np.where(data == data, np.where(data != 0, 2, 1), 2).astype(np.uint8)
            if k == 1:
                st[::5] |= 4
            assert np.array_equal(np.asarray(out.status, np.uint8), st.reshape(dims).transpose(perm).reshape(-1)), (dims, perm, k)
    assert taken >= 2


def test_large_stores_own_their_allocation(monkeypatch):
    """Stores of one batched call share one device allocation only while they are small: above the threshold
    (32 MiB per store; lowered to 0 here) every result owns its block, so destroying some measures of a cube
    returns their memory while the others live on.  The kept store must stay intact after its siblings are gone
    and after their blocks were reused."""
    import gc

    from olap_in_memory_b200 import interop

    G = _gpu()
    monkeypatch.setenv("OLAP_OWN_ARENA_MB", "0")
    rng = np.random.default_rng(51)
    n = 64 * 40 * 64
    datas = [cases.make_data(rng, n, 0.0, 0.7, "int") for _ in range(3)]
    stores = []
    for d in datas:
        s = G(n, "float32", 0)
        s.set_data_f32(d)
        stores.append(s)
    m = cases.random_map(rng, 40, 5, True)
    ident = cases.identity
    outs = G.drillUp_lowered(stores, [64, 40, 64], [64, 5, 64], [ident(64), m, ident(64)], ["sum", "highest", "first"])
    ptrs = [interop.values_tensor(o).data_ptr() for o in outs]
    plane = ((64 * 5 * 64 * 4 + 255) // 256) * 256
    assert ptrs[1] - ptrs[0] != plane or ptrs[2] - ptrs[1] != plane  # not carved from one block
    kept = outs[1].data_f32().copy()
    del outs[2], outs[0]
    gc.collect()
    scratch = [G(64 * 5 * 64, "float32", 0) for _ in range(4)]  # reuses the freed blocks
    for s in scratch:
        s.set_data_f32(np.full(64 * 5 * 64, 7.0, np.float32))
    assert np.array_equal(outs[0].data_f32().view(np.uint32), kept.view(np.uint32))
