"""olap_drill_up_pull on ONE GPU: the W ranks of a sharded cube are emulated as W sets of local
stores (the profiling guide forbids multi-rank kernels that wait on one another on one GPU; this
kernel never waits, its "peers" are simply other local allocations).  Every emulated rank pulls its
own output rows; the concatenation must equal the unsharded olap_drill_up BIT FOR BIT, values and
status, for every method — that is the claim the pull model makes over the push / NCCL exchange
(whose float32 partials only meet the 1e-6 tolerance)."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

METHODS = ["sum", "average", "highest", "lowest", "first", "last", "product"]


def _data(rng, n, default, kind):
    if kind == "dense":
        v = rng.uniform(1.0, 1000.0, n)
    elif kind == "mixed-sign":  # cancellation: float32-rounded partial sums would not survive this
        v = rng.uniform(-1e6, 1e6, n) * rng.choice([1.0, 1e-6], n)
    else:  # sparse
        v = rng.uniform(-5.0, 5.0, n)
        v[rng.random(n) < 0.6] = default
    v = v.astype(np.float32)
    if kind == "poison" or kind == "sparse":
        pass
    return v


def _pull_all_ranks(lib, N, GpuStore, full, rows_in, inner, in_bounds, out_bounds, full_map, methods, default, with_status, derive=False):
    from olap_in_memory_b200.sharded import _pull_tables
    from olap_in_memory_b200.store import _method_code

    W, K = len(in_bounds) - 1, len(methods)
    shards = []  # [rank][k]
    for r in range(W):
        lo, hi = in_bounds[r] * inner, in_bounds[r + 1] * inner
        per = []
        for k in range(K):
            s = GpuStore(hi - lo, "float32", default, with_status=with_status, shareable=(r % 2 == 0))
            s.set_data_f32(full[k][lo:hi])
            per.append(s)
        shards.append(per)
    base_v = (C.c_void_p * (K * W))(*[lib.olap_store_values_cptr(shards[r][k]._h) for k in range(K) for r in range(W)])
    base_s = (C.c_void_p * (K * W))(*[lib.olap_store_status_cptr(shards[r][k]._h) for k in range(K) for r in range(W)]) if with_status and not derive else None
    if derive:  # stores filled by set_data_f32: their status planes follow from the values, the kernel must not need them
        assert all(s.status_derived for per in shards for s in per)
    rank_rows = N.i64_array([in_bounds[r + 1] - in_bounds[r] for r in range(W)])
    values, status = [[] for _ in range(K)], [[] for _ in range(K)]
    for me in range(W):
        j0, j1 = out_bounds[me], out_bounds[me + 1]
        row_start, child_rank, child_row = _pull_tables(full_map, in_bounds, j0, j1)
        out = (C.c_void_p * K)()
        N.check(lib.olap_drill_up_pull(N.store_array([s._h for s in shards[me]]), K, N.int_array([_method_code(m) for m in methods]),
                                       j1 - j0, inner, row_start.ctypes.data_as(N.p_i32), child_rank.ctypes.data_as(N.p_i32),
                                       child_row.ctypes.data_as(N.p_i64), W, rank_rows, base_v, base_s, int(derive), out))
        assert lib.olap_last_op_path() == b"drillup/pull-peers"
        for k in range(K):
            st = GpuStore._wrap(out[k])
            values[k].append(st.data_f32())
            status[k].append(np.asarray(st.status, dtype=np.uint8))
    return [np.concatenate(v) for v in values], [np.concatenate(s) for s in status]


@pytest.mark.parametrize("default", [0.0, math.nan])
@pytest.mark.parametrize("with_status", [True, False])
@pytest.mark.parametrize("inner", [64, 36, 7])  # 128-bit path, 128-bit with a ragged last block, scalar path
def test_pull_equals_the_unsharded_rollup(default, with_status, inner):
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200.store import GpuStore

    N.init(0)
    lib = N.lib()
    rng = np.random.default_rng(5)
    d0, d1 = 7, 5  # rows = (d0, d1) flattened, d0 rolls up to 3 groups -> output rows (3, d1)
    rows_in, groups = d0 * d1, 3
    gmap = np.array([0, 2, 1, 0, 0, 2, 1], dtype=np.int32)
    full_map = (gmap[np.arange(rows_in) // d1] * d1 + np.arange(rows_in) % d1).astype(np.int64)
    rows_out = groups * d1
    for kind in ("dense", "sparse", "mixed-sign"):
        full = [_data(rng, rows_in * inner, default, kind) for _ in METHODS]
        if kind == "sparse" and default != default:
            full[0][:inner] = np.inf  # inf + -inf among the children of one parent: the exact-redo corner
            full[0][3 * d1 * inner:3 * d1 * inner + inner] = -np.inf
        for in_bounds, out_bounds in (([0, 9, 18, 27, 35], [0, 4, 8, 12, 15]), ([0, 0, 20, 20, 35], [0, 15, 15, 15, 15]),
                                      ([0, 35], [0, 15]), ([0, 1, 2, 3, 4, 5, 6, 35], [0, 3, 3, 6, 9, 9, 12, 15])):
            got_v, got_s = _pull_all_ranks(lib, N, GpuStore, full, rows_in, inner, in_bounds, out_bounds, full_map, METHODS, default, with_status)
            if with_status:  # the same without ever reading a status plane (derived from the values)
                der_v, der_s = _pull_all_ranks(lib, N, GpuStore, full, rows_in, inner, in_bounds, out_bounds, full_map, METHODS, default, True, derive=True)
                for k in range(len(METHODS)):
                    assert np.array_equal(der_v[k].view(np.uint32), got_v[k].view(np.uint32)) and np.array_equal(der_s[k], got_s[k])
            whole = []
            for k in range(len(METHODS)):
                s = GpuStore(rows_in * inner, "float32", default, with_status=with_status)
                s.set_data_f32(full[k])
                whole.append(s)
            ref = GpuStore.drillUp_lowered(whole, [d0, d1, inner], [groups, d1, inner], [gmap, None, None], METHODS)
            for k, method in enumerate(METHODS):
                want = ref[k].data_f32()
                assert got_v[k].shape == want.shape == (rows_out * inner,)
                assert np.array_equal(got_v[k].view(np.uint32), want.view(np.uint32)), (kind, method, in_bounds)
                assert np.array_equal(got_s[k], np.asarray(ref[k].status, dtype=np.uint8)), (kind, method, in_bounds)


def test_pull_with_a_shared_status_plane_and_argument_errors():
    from olap_in_memory_b200 import _native as N
    from olap_in_memory_b200.store import GpuStore

    N.init(0)
    lib = N.lib()
    n, inner = 6, 16
    out = (C.c_void_p * 2)()
    types, kinds = N.int_array([2, 2]), N.int_array([0, 0])
    N.check(lib.olap_store_create_batch(2, n * inner, types, kinds, 1 | 4, 1, out))  # status plane shared, shareable
    a, b = GpuStore._wrap(out[0]), GpuStore._wrap(out[1])
    assert lib.olap_store_status_ptr(a._h) == lib.olap_store_status_ptr(b._h)
    rng = np.random.default_rng(1)
    va = rng.uniform(1, 9, n * inner).astype(np.float32)
    a.set_data_f32(va)
    b.set_data_f32(va * 2)
    handle, v_off, s_off = a.ipc_export()
    assert len(handle) == 64 and v_off == 0 and s_off > 0
    row_start = np.array([0, 3, 6], dtype=np.int32)
    child_rank = np.zeros(6, dtype=np.int32)
    child_row = np.array([0, 2, 4, 1, 3, 5], dtype=np.int64)
    base_v = (C.c_void_p * 2)(lib.olap_store_values_cptr(a._h), lib.olap_store_values_cptr(b._h))
    base_s = (C.c_void_p * 2)(lib.olap_store_status_cptr(a._h), lib.olap_store_status_cptr(b._h))
    assert not a.status_derived  # a plane shared by several stores is never "derived"
    res = (C.c_void_p * 2)()
    args = lambda rows: (N.store_array([a._h, b._h]), 2, N.int_array([0, 5]), 2, inner, row_start.ctypes.data_as(N.p_i32),
                         child_rank.ctypes.data_as(N.p_i32), rows.ctypes.data_as(N.p_i64), 1, N.i64_array([n]), base_v, base_s, 0, res)
    N.check(lib.olap_drill_up_pull(*args(child_row)))
    ra, rb = GpuStore._wrap(res[0]), GpuStore._wrap(res[1])
    assert lib.olap_store_status_ptr(ra._h) == lib.olap_store_status_ptr(rb._h)  # the results share a plane too
    m = va.reshape(n, inner).astype(np.float64)
    assert np.array_equal(ra.data_f32(), np.stack([m[0] + m[2] + m[4], m[1] + m[3] + m[5]]).astype(np.float32).ravel())
    assert np.array_equal(rb.data_f32(), (va.reshape(n, inner)[[4, 5]] * 2).ravel())  # last
    assert set(ra.status) == {2}
    # results of a shareable store are shareable (the next rollup may pull from them)
    assert len(ra.ipc_export()[0]) == 64
    plain = GpuStore(4, "float32", 0.0)
    with pytest.raises(N.OlapError, match="OLAP_CREATE_SHAREABLE"):
        plain.ipc_export()
    bad = child_row.copy()
    bad[2] = n  # a child row the rank does not hold
    with pytest.raises(N.OlapValueError, match="holds 6 rows"):
        N.check(lib.olap_drill_up_pull(*args(bad)))
