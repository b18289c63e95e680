"""BASELINE.json's full sizes, where no CPU oracle finishes in seconds: size-independent
properties through the C ABI (checksums of checksums, order relations, round trips,
idempotence), plus a sampled exact comparison against a float64 numpy reference."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GEN = 32
C, P = 3652, 120
I = GEN ** 3  # config 2: 3652 x 32^3 = 119 668 736 cells


def _setup():
    import torch

    from olap_in_memory_b200 import TimeDimension, _native, interop
    from olap_in_memory_b200.store import GpuStore

    _native.init(0)
    interop.use_torch_stream()
    day = TimeDimension("time", "day", "2010-01-01", "2019-12-31")
    month = np.asarray(day.getGroupIndexFromRootIndexMap("month"), np.int32)
    return torch, interop, GpuStore, month


def _filled(torch, interop, GpuStore, n, default, fill, seed):
    s = GpuStore(n, "float32", default)
    v = interop.values_tensor(s)
    g = torch.Generator(device="cuda").manual_seed(seed)
    v.copy_(torch.randint(1, 1000, (n,), generator=g, device="cuda").float())
    if fill < 1.0:
        drop = torch.rand(n, generator=g, device="cuda") >= fill
        v[drop] = math.nan if default != default else 0.0
    st = interop.status_tensor(s)
    if st is not None:
        pres = (v == v) if default != default else (v != 0)
        st.copy_(torch.where(pres, 2, 1).to(torch.uint8))
    torch.cuda.synchronize()
    return s


@pytest.mark.parametrize("default,fill", [(0.0, 1.0), (math.nan, 0.5)])
def test_config2_drillup_properties(default, fill):
    torch, interop, G, month = _setup()
    n = C * I
    ident = np.arange(I, dtype=np.int32)
    src = _filled(torch, interop, G, n, default, fill, 1)
    methods = ["sum", "average", "highest", "lowest", "first", "last"]
    outs = G.drillUp_lowered([src] * len(methods), [C, I], [P, I], [month, ident], methods)
    r = {m: interop.values_tensor(o).view(P, I) for m, o in zip(methods, outs)}
    x = interop.values_tensor(src).view(C, I)
    pres = (x == x) if default != default else (x != 0)
    # checksum of checksums: integers < 1000, so every sum is exact in float32/float64
    assert math.isclose(outs[0].total, src.total, rel_tol=1e-12)
    # order relations wherever a parent has a set child
    has = (r["sum"] == r["sum"]) if default != default else (r["highest"] != 0)
    assert bool(torch.all((r["lowest"] <= r["average"])[has])) and bool(torch.all((r["average"] <= r["highest"])[has]))
    # exact check of whole columns against float64 torch on the device data (sampled columns)
    cols = torch.tensor([0, 1, 12345, I // 2, I - 1], device="cuda")
    mm = torch.from_numpy(month.astype(np.int64)).cuda()
    xs = x[:, cols].double()
    ps = pres[:, cols]
    want_sum = torch.zeros(P, cols.numel(), dtype=torch.float64, device="cuda").index_add_(0, mm, torch.where(ps, xs, 0.0))
    want_cnt = torch.zeros(P, cols.numel(), dtype=torch.float64, device="cuda").index_add_(0, mm, ps.double())
    got_sum = r["sum"][:, cols].double()
    none = want_cnt == 0
    assert bool(torch.all(torch.where(none, True, got_sum == want_sum)))
    if default != default:
        assert bool(torch.all(torch.isnan(got_sum[none])))
    got_avg = r["average"][:, cols]
    want_avg = (want_sum / want_cnt).float()
    assert bool(torch.all(torch.where(none, True, got_avg == want_avg)))
    # first / last = first / last SET child of each month
    for col in range(cols.numel()):
        xc, pc = xs[:, col], ps[:, col]
        for p_ in (0, 57, 119):
            idx = torch.nonzero((mm == p_) & pc).flatten()
            if idx.numel():
                assert float(r["first"][p_, cols[col]]) == float(xc[idx[0]])
                assert float(r["last"][p_, cols[col]]) == float(xc[idx[-1]])
    # status: a month is complete (0x2) only if all its days are set
    st = interop.status_tensor(outs[0])
    if st is not None:
        st = st.view(P, I)[:, cols]
        full = want_cnt == torch.bincount(mm, minlength=P).double().unsqueeze(1)
        assert bool(torch.all(st[full] == 2)) and bool(torch.all((st[~full & ~none] == 3))) and bool(torch.all(st[none] == 1))
    # the tile kernel (time innermost) must agree with the mid kernel on the transposed cube
    small_i = 4096
    sub = G(C * small_i, "float32", default)
    interop.values_tensor(sub).view(C, small_i).copy_(x[:, :small_i])
    stt = interop.status_tensor(sub)
    if stt is not None:
        stt.view(C, small_i).copy_(interop.status_tensor(src).view(C, I)[:, :small_i])
    torch.cuda.synchronize()
    tr = G.reorder_lowered([sub], [C, small_i], [1, 0])[0]
    a = G.drillUp_lowered([tr], [small_i, C], [small_i, P], [np.arange(small_i, dtype=np.int32), month], ["sum"])[0]
    b = G.drillUp_lowered([sub], [C, small_i], [P, small_i], [month, np.arange(small_i, dtype=np.int32)], ["sum"])[0]
    ta, tb = interop.values_tensor(a).view(small_i, P), interop.values_tensor(b).view(P, small_i)
    assert bool(torch.all((ta.t() == tb) | (torch.isnan(ta.t()) & torch.isnan(tb))))


def test_config3_style_round_trips():
    """dice -> reorder -> inverse reorder is the identity on the kept cells; drillDown then
    drillUp restores the parents (float32 'sum' spreading, exact for these values)."""
    torch, interop, G, _ = _setup()
    dims = [50, 100, 100, 10, 10, 10]  # 5e8 cells
    n = int(np.prod(dims))
    src = _filled(torch, interop, G, n, 0.0, 0.7, 3)
    ident = [np.arange(d, dtype=np.int32) for d in dims]
    keep = list(ident)
    keep[0] = np.arange(0, 50, 2, dtype=np.int32)
    diced = G.dice_lowered([src], dims, keep)[0]
    ddims = [25] + dims[1:]
    x = interop.values_tensor(src).view(*dims)
    assert bool(torch.equal(interop.values_tensor(diced).view(*ddims), x[::2]))
    perm = [5, 4, 3, 2, 1, 0]
    rev = G.reorder_lowered([diced], ddims, perm)[0]
    assert bool(torch.equal(interop.values_tensor(rev).view(*ddims[::-1]), x[::2].permute(*perm)))
    back = G.reorder_lowered([rev], ddims[::-1], perm)[0]
    assert bool(torch.equal(interop.values_tensor(back), interop.values_tensor(diced)))
    assert interop.status_tensor(back) is None or bool(torch.equal(interop.status_tensor(back), interop.status_tensor(diced)))
    del rev, back, diced
    # drillDown axis 3 (10 items -> 40, 4 children each) then drillUp back
    down_map = np.repeat(np.arange(10, dtype=np.int32), 4)
    small = [20, 100, 100, 10, 10, 10]
    s2 = _filled(torch, interop, G, int(np.prod(small)), 0.0, 0.7, 4)
    interop.values_tensor(s2).mul_(4.0)  # divisible by the 4 children: v/4 exact
    torch.cuda.synchronize()
    maps = [np.arange(d, dtype=np.int32) for d in small]
    maps[3] = down_map
    big = list(small)
    big[3] = 40
    down = G.drillDown_lowered([s2], small, big, maps, ["sum"])[0]
    up_maps = [np.arange(d, dtype=np.int32) for d in big]
    up_maps[3] = down_map
    up = G.drillUp_lowered([down], big, small, up_maps, ["sum"])[0]
    assert bool(torch.equal(interop.values_tensor(up), interop.values_tensor(s2)))
    st = interop.status_tensor(up)
    if st is not None:
        pres = interop.values_tensor(s2) != 0
        assert bool(torch.all(st[pres] == 6)) and bool(torch.all(st[~pres] == 1))  # set + interpolated


def test_more_than_2_31_cells():
    """Linear indices are int64 (SURVEY 'hard parts'): a 2.4e9-cell store through the mid
    drillUp kernel, the 64-bit scalar gather (reorder that keeps a short inner run) and the
    tile kernel with > 2^31 input cells; checked by totals, order relations and sampled cells."""
    torch, interop, G, _ = _setup()
    old = G.WITH_STATUS
    G.WITH_STATUS = False  # 4 B/cell keeps the case at ~10 GB
    try:
        A, B, C = 3, 40000, 20000  # 2.4e9 cells
        n = A * B * C
        assert n > 2 ** 31
        src = G(n, "float32", 0.0)
        v = interop.values_tensor(src)
        g = torch.Generator(device="cuda").manual_seed(11)
        chunk = 1 << 28
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            v[lo:hi] = torch.randint(1, 100, (hi - lo,), generator=g, device="cuda").float()
        torch.cuda.synchronize()
        x = v.view(A, B, C)
        # mid kernel: axis 1, 40000 -> 4 groups of 10000
        m = (np.arange(B) // 10000).astype(np.int32)
        ident = lambda k: np.arange(k, dtype=np.int32)
        up = G.drillUp_lowered([src], [A, B, C], [A, 4, C], [ident(A), m, ident(C)], ["sum"])[0]
        r = interop.values_tensor(up).view(A, 4, C)
        cols = torch.tensor([0, 7777, C - 1], device="cuda")
        want = x[:, :, cols].double().view(A, 4, 10000, 3).sum(dim=2)
        assert bool(torch.equal(r[:, :, cols].double(), want))
        assert math.isclose(up.total, src.total, rel_tol=1e-12)
        # the last cells of the store (offsets beyond 2^31) are really addressed
        assert float(r[A - 1, 3, C - 1]) == float(x[A - 1, 30000:, C - 1].double().sum())
        del up, r
        # tile kernel: innermost axis 20000 -> 2 (rows of 20000 cells, 1.2e5 rows)
        m2 = (np.arange(C) // 10000).astype(np.int32)
        up2 = G.drillUp_lowered([src], [A * 4, B // 4, C], [A * 4, B // 4, 2],
                                [ident(A * 4), ident(B // 4), m2], ["highest"])[0]
        r2 = interop.values_tensor(up2).view(A, B, 2)
        rows = torch.tensor([0, 12345, B - 1], device="cuda")
        want2 = x[:, rows, :].view(A, 3, 2, 10000).amax(dim=3)
        assert bool(torch.equal(r2[:, rows, :], want2))
        del up2, r2
        # scalar gather with 64-bit offsets: view as [A, B, K, 5] and swap the middle axes
        # (inner run of 5 cells, 4.8e8 rows, source offsets beyond 2^31)
        K = C // 5
        re = G.reorder_lowered([src], [A, B, K, 5], [0, 2, 1, 3])[0]
        from olap_in_memory_b200 import _native

        assert _native.lib().olap_last_op_path().decode() in ("gather/scalar-big", "gather/rows")
        y = interop.values_tensor(re).view(A, K, B, 5)
        x4 = x.view(A, B, K, 5)
        for k_ in (0, 1234, K - 1):
            assert bool(torch.equal(y[:, k_], x4[:, :, k_]))
    finally:
        G.WITH_STATUS = old


@pytest.mark.parametrize("O,Cn,I_,Pn", [(1, 100_000_000, 1, 1), (1500, 45_000, 1, 3), (4, 2_000_000, 5, 2)])
def test_long_rows_against_torch(O, Cn, I_, Pn):
    """drillup/long (rows too long for a tile, few parents): a 1e8-cell collapse, the
    one-CTA-per-row regime without merge pass, and a short inner run — against torch."""
    from olap_in_memory_b200 import _native as N

    torch, interop, G, _ = _setup()
    n = O * Cn * I_
    src = _filled(torch, interop, G, n, 0.0, 0.6, 11)
    cut = np.sort(np.random.default_rng(3).integers(1, Cn, Pn - 1)) if Pn > 1 else np.array([], np.int64)
    m = np.searchsorted(cut, np.arange(Cn), side="right").astype(np.int32)
    ident = lambda k: np.arange(k, dtype=np.int32)
    methods = ["sum", "highest", "first", "last", "average"]
    outs = G.drillUp_lowered([src] * len(methods), [O, Cn, I_], [O, Pn, I_], [ident(O), m, ident(I_)], methods)
    # many long rows with few parents and no inner run go to the lanes kernel when no status plane has to be read
    assert N.lib().olap_last_op_path().decode() in ("drillup/lanes", "drillup/long")
    x = interop.values_tensor(src).view(O, Cn, I_)
    bounds = [0] + cut.tolist() + [Cn]
    for p_ in range(Pn):
        seg = x[:, bounds[p_]:bounds[p_ + 1], :]
        pres = seg != 0
        want_sum = seg.double().sum(1)
        cnt = pres.sum(1)
        got = {k: interop.values_tensor(o).view(O, Pn, I_)[:, p_, :] for k, o in zip(methods, outs)}
        assert bool(torch.equal(got["sum"], want_sum.float()))
        assert bool(torch.equal(got["highest"], seg.max(1).values))  # values are >= 0, unset cells are 0
        avg = torch.where(cnt > 0, want_sum / cnt.clamp(min=1).double(), torch.zeros_like(want_sum)).float()
        assert bool(torch.equal(got["average"], avg))
        idx = torch.arange(seg.shape[1], device="cuda").view(1, -1, 1)
        first_i = torch.where(pres, idx, seg.shape[1]).min(1).values.clamp(max=seg.shape[1] - 1)
        last_i = torch.where(pres, idx, -1).max(1).values.clamp(min=0)
        assert bool(torch.equal(got["first"], torch.gather(seg, 1, first_i.unsqueeze(1)).squeeze(1)))
        assert bool(torch.equal(got["last"], torch.gather(seg, 1, last_i.unsqueeze(1)).squeeze(1)))
        st = interop.status_tensor(outs[0])
        if st is not None:
            sp = st.view(O, Pn, I_)[:, p_, :]
            full = cnt == seg.shape[1]
            assert bool(torch.all(sp[full] == 2)) and bool(torch.all(sp[(cnt > 0) & ~full] == 3)) and bool(torch.all(sp[cnt == 0] == 1))
