"""The multi-GPU path on CPU: world_size 2 (and 3) over gloo, with the CPU oracle standing
in for the device store, must reproduce what ONE cube does — local transforms without
communication, and drillUp of a sharded dimension through partial rollup + all-to-all +
ordered combine (olap_in_memory_b200/sharded.py)."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
METHODS = ["sum", "average", "highest", "lowest", "first", "last"]
ROLLUPS = (("region", "country"), ("region", "all"), ("product", "family"), ("time", "quarter"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dims():
    from olap_in_memory_b200 import GenericDimension, TimeDimension

    region = GenericDimension("region", "city", [f"c{i}" for i in range(7)])
    region.addAttribute("city", "country", lambda c: "even" if int(c[1:]) % 2 == 0 else "odd")
    product = GenericDimension("product", "sku", [f"p{i}" for i in range(5)])
    product.addAttribute("sku", "family", lambda p: f"f{int(p[1:]) // 2}")
    time = TimeDimension("time", "month", "2010-01", "2010-12")
    return [region, product, time]


def _data(default, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases

    rng = np.random.default_rng(seed)
    return cases.make_data(rng, 7 * 5 * 12, default, 0.6, "int").astype(np.float64)


def _fill(cube, default):
    for k, method in enumerate(METHODS):
        cube.createStoredMeasure(f"m_{method}", {"region": method, "product": method, "time": method}, "float32", default)
        cube.setData(f"m_{method}", _data(default, k).tolist())


def _collect(cube, ids, extended=True):
    """`extended`: also the transforms that move whole rows between ranks on request (rebalance, permuting dice of a
    sharded dimension).  tests/gpu_sharded_check.py switches them on with OLAP_SHARDED_EXTENDED=1: they are verified
    over gloo here, but were written after the round's last multi-GPU hardware call."""
    results = {}
    for dim, attr in ROLLUPS:
        rolled = cube.drillUp(dim, attr)
        for m in ids:
            results[(dim, attr, m)] = np.asarray(rolled.getData(m), dtype=np.float64)
    chained = cube.drillUp("time", "quarter").drillUp("region", "all").drillUp("product", "all")
    for m in ids:
        results[("chain", "all", m)] = np.asarray(chained.getData(m), dtype=np.float64)
    results["total"] = cube.getTotal("m_sum")
    diced = cube.dice("time", "month", ["2010-03", "2010-04", "2010-05"])
    results["dice"] = np.asarray(diced.getData("m_sum"), dtype=np.float64)
    # dice of the (possibly) sharded dimensions: rows are dropped where they live, the shards
    # become uneven, and a rollup of the diced dimension still has to come out right
    diced = cube.dice("region", "city", ["c1", "c2", "c5", "c6"])
    results["dice_region"] = np.asarray(diced.getData("m_sum"), dtype=np.float64)
    results["dice_region_total"] = diced.getTotal("m_sum")
    for m in ids:
        results[("dice_region", "country", m)] = np.asarray(diced.drillUp("region", "country").getData(m), dtype=np.float64)
    sparse = cube.dice("product", "sku", ["p0", "p3"]).dice("region", "country", ["odd"])
    for m in ids:
        results[("dice_product_region", "all", m)] = np.asarray(sparse.drillUp("product", "all").getData(m), dtype=np.float64)
    results["dice_one_row"] = np.asarray(cube.dice("region", "city", ["c6"]).drillUp("region", "all").getData("m_last"), dtype=np.float64)
    if not extended:
        return results
    # ... and after the rows were spread evenly again (ShardedCube.rebalance: one all-to-all per plane)
    even = diced
    if hasattr(diced, "rebalance"):
        from olap_in_memory_b200.sharded import split_rows

        even = diced.rebalance()
        assert list(even.row_bounds) == list(split_rows(even.rows_total, even.world)), (even.row_bounds, diced.row_bounds)
        assert even.rebalance() is even
    results["rebalanced"] = np.asarray(even.getData("m_first"), dtype=np.float64)
    for m in ids:
        results[("rebalanced", "country", m)] = np.asarray(even.drillUp("region", "country").getData(m), dtype=np.float64)
        results[("rebalanced", "quarter", m)] = np.asarray(even.drillUp("time", "quarter").getData(m), dtype=np.float64)
    # dice(reorder=True) that permutes the (possibly) sharded dimension: rows dropped in place, then shuffled between ranks
    for items in (["c5", "c1", "c6", "c2"], ["c6", "c5", "c4", "c3", "c2", "c1", "c0"], ["c3", "c0"]):
        shuffled = cube.dice("region", "city", items, True)
        assert shuffled.getDimension("region").getItems() == items if hasattr(shuffled, "getDimension") else True
        results[("dice_permuted", tuple(items))] = np.asarray(shuffled.getData("m_last"), dtype=np.float64)
        # (first / last of a rollup AFTER a permuting dice follow the Map's insertion order in the reference, SURVEY.md
        # A13 / A14: a declared divergence, not compared here)
        results[("dice_permuted_up", tuple(items))] = np.asarray(shuffled.drillUp("region", "country").getData("m_highest"), dtype=np.float64)
        results[("dice_permuted_sum", tuple(items))] = np.asarray(shuffled.drillUp("region", "all").getData("m_sum"), dtype=np.float64)
    permuted_product = cube.dice("product", "sku", ["p4", "p0", "p2"], True)
    results["dice_permuted_product"] = np.asarray(permuted_product.getData("m_sum"), dtype=np.float64)
    one = cube.dice("region", "city", ["c6"])  # every surviving row on the last rank
    one = one.rebalance() if hasattr(one, "rebalance") else one
    results["rebalanced_one_row"] = np.asarray(one.drillUp("product", "family").getData("m_average"), dtype=np.float64)
    return results


def _worker(rank, world, port, prefix, default_is_nan, queue):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from olap_in_memory_b200.sharded import ShardedCube
        from oracle.store_oracle import OracleStore
        from shard_store import OracleShardStore

        default = math.nan if default_is_nan else 0.0
        cube = ShardedCube(_dims(), prefix=prefix, store_cls=OracleShardStore)
        _fill(cube, default)
        results = _collect(cube, list(cube.storedMeasures))
        if rank == 0:
            queue.put(results)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _expected(default_is_nan):
    from olap_in_memory_b200 import Cube
    from oracle.store_oracle import OracleStore
    from shard_store import OracleShardStore

    cube = Cube(_dims(), OracleStore)
    _fill(cube, math.nan if default_is_nan else 0.0)
    return _collect(cube, cube.storedMeasureIds)


@pytest.mark.parametrize("world,prefix,default_is_nan", [(2, 1, False), (2, 2, True), (3, 2, False)])
def test_sharded_cube_matches_single_cube(world, prefix, default_is_nan):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, prefix, default_is_nan, queue)) for r in range(world)]
    for p in procs:
        p.start()
    got = queue.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _expected(default_is_nan)
    assert got.keys() == want.keys()
    for key in want:
        if key == "total":
            assert math.isclose(got[key], want[key], rel_tol=1e-12) or (math.isnan(got[key]) and math.isnan(want[key]))
            continue
        g, w = np.asarray(got[key], np.float64), np.asarray(want[key], np.float64)
        assert g.shape == w.shape, key
        # chained first/last after a sparse rollup is the documented F6-ii divergence of the
        # reference's Map order; the sharded cube follows index order
        if key[0] == "chain" and key[2] in ("m_first", "m_last"):
            continue
        assert np.allclose(g, w, rtol=1e-12, atol=0, equal_nan=True), (key, g[:8], w[:8])


def test_split_rows_is_balanced_and_contiguous():
    from olap_in_memory_b200.sharded import split_rows

    assert split_rows(100, 8) == [0, 13, 26, 39, 52, 64, 76, 88, 100]  # 13,13,13,13,12,12,12,12
    assert split_rows(10, 8) == [0, 2, 4, 5, 6, 7, 8, 9, 10]
    assert split_rows(3, 4) == [0, 1, 2, 3, 3]


# ---- drillDown of a sharded dimension (time outermost) -------------------------------------------
def _time_first_dims():
    from olap_in_memory_b200 import GenericDimension, TimeDimension

    region = GenericDimension("region", "city", [f"c{i}" for i in range(7)])
    return [TimeDimension("time", "quarter", "2010-Q1", "2011-Q4"), region]


def _time_first_collect(cube, types):
    out = {}
    down = cube.drillDown("time", "month")
    for m in types:
        out[("down", m)] = np.asarray(down.getData(m), dtype=np.float64)
        out[("down_up_year", m)] = np.asarray(down.drillUp("time", "year").getData(m), dtype=np.float64)
        out[("down_dice", m)] = np.asarray(down.dice("time", "month", ["2010-11", "2010-12", "2011-01", "2011-02"]).getData(m), dtype=np.float64)
    return out


def _time_first_fill(cube, default):
    rng = np.random.default_rng(5)
    for m, (type_, rule) in {"f_sum": ("float32", "sum"), "u_sum": ("uint32", "sum"), "f_avg": ("float32", "average")}.items():
        cube.createStoredMeasure(m, {"time": rule, "region": "sum"}, type_, default)
        values = rng.integers(0, 50, 8 * 7).astype(np.float64)
        values[rng.random(8 * 7) < 0.3] = default
        cube.setData(m, values.tolist())
    return ["f_sum", "u_sum", "f_avg"]


def _time_first_worker(rank, world, port, prefix, queue):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from olap_in_memory_b200.sharded import ShardedCube
        from oracle.store_oracle import OracleStore
        from shard_store import OracleShardStore

        cube = ShardedCube(_time_first_dims(), prefix=prefix, store_cls=OracleShardStore)
        ids = _time_first_fill(cube, 0.0)
        try:
            results = _time_first_collect(cube, ids)
        except NotImplementedError as e:
            results = {"unsupported": str(e)}
        if rank == 0:
            queue.put(results)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,prefix,supported", [(2, 1, True), (3, 1, True), (2, 2, True), (3, 2, True)])
def test_sharded_drilldown_of_the_sharded_dimension(world, prefix, supported):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_time_first_worker, args=(r, world, port, prefix, queue)) for r in range(world)]
    for p in procs:
        p.start()
    got = queue.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # (3, 2): 56 rows over 3 ranks, the bounds cut through a quarter: whole rows move first (_repartition), then the expansion
    if not supported:
        assert "re-partitions rows" in got["unsupported"]
        return
    from olap_in_memory_b200 import Cube
    from oracle.store_oracle import OracleStore
    from shard_store import OracleShardStore

    single = Cube(_time_first_dims(), OracleStore)
    want = _time_first_collect(single, _time_first_fill(single, 0.0))
    assert got.keys() == want.keys()
    for key in want:
        if key[0] == "down_up_year":  # partial sums per rank, then combined: association differs from one sequential sum
            assert np.allclose(got[key], want[key], rtol=1e-12, atol=0, equal_nan=True), key
            continue
        assert np.array_equal(np.asarray(got[key]), np.asarray(want[key]), equal_nan=True), (key, got[key][:12], want[key][:12])


# ---- reorderDimensions, including permutations that move the sharded dimension ---------------------
def _reorder_collect(cube, sharded_prefix=None):
    import itertools

    out = {}
    ids = ["m_sum", "m_first"]
    for order in itertools.permutations(["region", "product", "time"]):
        try:
            moved = cube.reorderDimensions(list(order))
        except NotImplementedError:
            assert sharded_prefix == 2  # only a cube sharded on (region, product) refuses, and says so
            out[order] = "unsupported"
            continue
        for m in ids:
            out[(order, m)] = np.asarray(moved.getData(m), dtype=np.float64)
        if order == ("product", "time", "region"):  # carry on working with the re-partitioned cube
            for m in ids:
                out[(order, "family", m)] = np.asarray(moved.drillUp("product", "family").getData(m), dtype=np.float64)
                out[(order, "diced", m)] = np.asarray(moved.dice("product", "sku", ["p1", "p4"]).getData(m), dtype=np.float64)
                out[(order, "back", m)] = np.asarray(moved.reorderDimensions(["region", "product", "time"]).getData(m),
                                                     dtype=np.float64)
    # a rollup of the sharded dimension that leaves fewer rows than ranks shards the result on the next dimension too
    if sharded_prefix == 1:
        rolled = cube.drillUp("region", "all")
        assert rolled.prefix == 2 and rolled.rows_total == 5
        sizes = np.diff(rolled.row_bounds)
        assert sizes.max() - sizes.min() <= 1 and sizes.sum() == 5, rolled.row_bounds
        assert cube.drillUp("region", "country").prefix == (1 if cube.world <= 2 else 2)
    # compositions: diceRange, slice, removeDimension(s), keepDimensions on sharded and inner dimensions
    ranged = cube.diceRange("time", "month", "2010-03", "2010-07")
    out["diceRange"] = np.asarray(ranged.getData("m_sum"), dtype=np.float64)
    for m in ids:
        out[("remove_region", m)] = np.asarray(cube.removeDimension("region").getData(m), dtype=np.float64)
        out[("remove_time", m)] = np.asarray(cube.removeDimension("time").getData(m), dtype=np.float64)
        out[("slice_region", m)] = np.asarray(cube.slice("region", "city", "c3").getData(m), dtype=np.float64)
        out[("slice_country", m)] = np.asarray(cube.slice("region", "country", "odd").getData(m), dtype=np.float64)
        if m != "m_first":  # chained first/last after a sparse rollup: declared divergence (Map order)
            out[("keep_time", m)] = np.asarray(cube.keepDimensions(["time"]).getData(m), dtype=np.float64)
    assert cube.removeDimension("region").dimensionIds == ["product", "time"]
    assert cube.keepDimensions(["time"]).dimensionIds == ["time"]
    # computed measures: shard-local evaluation, `__total` through one all-reduce, formulas carried through transforms
    cube.createComputedMeasure("ratio", "(m_sum + m_highest) / m_lowest")
    cube.createComputedMeasure("share", "m_sum / m_sum__total + ratio")
    out["ratio"] = np.asarray(cube.getData("ratio"), dtype=np.float64)
    out["share"] = np.asarray(cube.getData("share"), dtype=np.float64)
    out["share_rolled"] = np.asarray(cube.drillUp("region", "country").drillUp("time", "quarter").getData("share"), dtype=np.float64)
    # a dice that leaves the shards uneven (and one rank empty on 3 ranks), then the exchange
    diced = cube.dice("region", "city", ["c0", "c1", "c2"])
    out["share_diced"] = np.asarray(diced.getData("share"), dtype=np.float64)
    try:
        out["uneven"] = np.asarray(diced.reorderDimensions(["time", "region", "product"]).getData("m_sum"), dtype=np.float64)
    except NotImplementedError:
        assert sharded_prefix == 2
        out["uneven"] = "unsupported"
    return out


def _reorder_worker(rank, world, port, prefix, default_is_nan, queue):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from olap_in_memory_b200.sharded import ShardedCube
        from oracle.store_oracle import OracleStore
        from shard_store import OracleShardStore

        cube = ShardedCube(_dims(), prefix=prefix, store_cls=OracleShardStore)
        _fill(cube, math.nan if default_is_nan else 0.0)
        results = _reorder_collect(cube, prefix)
        if rank == 0:
            queue.put(results)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,prefix,default_is_nan", [(2, 1, False), (3, 1, True), (2, 2, False)])
def test_sharded_reorder_matches_single_cube(world, prefix, default_is_nan):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reorder_worker, args=(r, world, port, prefix, default_is_nan, queue)) for r in range(world)]
    for p in procs:
        p.start()
    got = queue.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from olap_in_memory_b200 import Cube
    from oracle.store_oracle import OracleStore
    from shard_store import OracleShardStore

    single = Cube(_dims(), OracleStore)
    _fill(single, math.nan if default_is_nan else 0.0)
    want = _reorder_collect(single)
    compared = 0
    for key, value in got.items():
        if isinstance(value, str):
            assert prefix == 2
            continue
        if isinstance(key, str) and key.startswith("share"):  # the total is a sum of per-rank sums: association differs
            assert np.allclose(value, want[key], rtol=1e-12, atol=0, equal_nan=True), key
        else:
            assert np.array_equal(value, want[key], equal_nan=True), key  # pure data movement / per-cell formulas: bit-exact
        compared += 1
    assert compared == 33


def test_reorder_inside_the_shard_is_local():
    """Single rank, four dimensions, cube sharded on the first two: permutations of the other two
    (and nothing else) are accepted and match one cube."""
    sys.path.insert(0, ROOT)
    from olap_in_memory_b200 import Cube, GenericDimension
    from olap_in_memory_b200.sharded import ShardedCube
    from oracle.store_oracle import OracleStore
    from shard_store import OracleShardStore

    def dims():
        return [GenericDimension(name, "root", [f"{name}{i}" for i in range(n)]) for name, n in (("a", 2), ("b", 3), ("c", 4), ("d", 5))]

    values = np.arange(1, 121, dtype=np.float64)
    values[::7] = 0
    sharded, single = ShardedCube(dims(), prefix=2, store_cls=OracleShardStore), Cube(dims(), OracleStore)
    for cube in (sharded, single):
        cube.createStoredMeasure("mm", {}, "float32", 0)
        cube.setData("mm", values.tolist())
    moved = sharded.reorderDimensions(["a", "b", "d", "c"])
    assert moved.dimensionIds == ["a", "b", "d", "c"]
    assert np.array_equal(moved.getData("mm"), np.asarray(single.reorderDimensions(["a", "b", "d", "c"]).getData("mm")))
    assert sharded.reorderDimensions(["a", "b", "c", "d"]) is sharded
    # a permutation that moves a sharded dimension: the cube is first sharded on its outermost dimension alone
    # (whole rows move between neighbouring ranks), then re-partitioned on the dimension that comes to the front
    for order in (["b", "a", "c", "d"], ["d", "c", "b", "a"], ["c", "a", "d", "b"]):
        moved = sharded.reorderDimensions(order)
        assert moved.dimensionIds == order and moved.prefix == 1
        assert np.array_equal(moved.getData("mm"), np.asarray(single.reorderDimensions(order).getData("mm")))
    with pytest.raises(ValueError):
        sharded.reorderDimensions(["a", "b", "c", "c"])


# ---- random operation sequences: the sharded cube must track ONE cube step by step -------------------
def _fuzz_dims(rng):
    from olap_in_memory_b200 import GenericDimension, TimeDimension

    region = GenericDimension("region", "city", [f"c{i}" for i in range(7)])
    region.addAttribute("city", "country", lambda c: "even" if int(c[1:]) % 2 == 0 else "odd")
    product = GenericDimension("product", "sku", [f"p{i}" for i in range(5)])
    product.addAttribute("sku", "family", lambda p: f"f{int(p[1:]) // 2}")
    dims = [region, product, TimeDimension("time", "month", "2010-01", "2010-12")]
    order = rng.permutation(3)
    return [dims[i] for i in order]


def _fuzz_op(rng, cube):
    """One random transform as (method name, args), chosen from the cube's current dimensions."""
    dims = cube.dimensions
    dim = dims[int(rng.integers(len(dims)))]
    kind = rng.choice(["drillUp", "dice", "reorder", "remove", "drillDown", "drillUp", "dice"])
    if kind == "drillUp":
        choices = [a for a in dim.attributes if a != dim.rootAttribute]
        if choices:
            return "drillUp", (dim.id, str(rng.choice(choices)))
    if kind == "dice" and dim.numItems > 1:
        items = dim.getItems()
        if dim.id == "time":
            lo = int(rng.integers(len(items)))
            hi = int(rng.integers(lo, len(items)))
            return "diceRange", (dim.id, dim.rootAttribute, items[lo], items[hi])
        keep = [it for it in items if rng.random() < 0.6] or [items[0]]
        return "dice", (dim.id, dim.rootAttribute, keep)
    if kind == "reorder" and len(dims) > 1:
        return "reorderDimensions", ([dims[i].id for i in rng.permutation(len(dims))],)
    if kind == "remove" and len(dims) > 1:
        return "removeDimension", (dim.id,)
    time = next((d for d in dims if d.id == "time"), None)
    if time is not None and time.rootAttribute in ("quarter", "semester", "year"):
        return "drillDown", ("time", "month")
    return "reorderDimensions", ([d.id for d in dims],)  # no-op


def _fuzz_worker(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from olap_in_memory_b200 import Cube
        from olap_in_memory_b200.sharded import ShardedCube
        from oracle.store_oracle import OracleStore
        from shard_store import OracleShardStore

        log = []
        for seed in range(int(os.environ.get("OLAP_FUZZ_SEEDS", "24"))):
            rng = np.random.default_rng(100 + seed)  # the same stream on every rank
            prefix = 1 + seed % 2
            default = math.nan if seed % 3 == 0 else 0.0
            dims = _fuzz_dims(rng)
            sharded, single = ShardedCube(dims, prefix=prefix, store_cls=OracleShardStore), Cube(dims, OracleStore)
            values = rng.integers(1, 100, 7 * 5 * 12).astype(np.float64)
            values[rng.random(values.size) < 0.3] = default
            for cube in (sharded, single):
                for m, rule in (("m_sum", "sum"), ("m_high", "highest")):
                    cube.createStoredMeasure(m, {"region": rule, "product": rule, "time": rule}, "float32", default)
                    cube.setData(m, values.tolist())
            for step in range(7):
                name, args = _fuzz_op(rng, single)
                try:
                    moved = getattr(sharded, name)(*args)
                except NotImplementedError:
                    log.append((seed, step, name, "refused"))
                    continue
                sharded, single = moved, getattr(single, name)(*args)
                if name in ("dice", "diceRange") and (seed + step) % 2 == 0:
                    sharded = sharded.rebalance()  # rows spread evenly again; the cube must stay the same cube
                    name = name + "+rebalance"
                assert sharded.dimensionIds == single.dimensionIds, (seed, step, name, args)
                for m in ("m_sum", "m_high"):
                    got, want = np.asarray(sharded.getData(m), np.float64), np.asarray(single.getData(m), np.float64)
                    ok = got.shape == want.shape and np.allclose(got, want, rtol=1e-12, atol=0, equal_nan=True)
                    log.append((seed, step, name, "ok" if ok else f"MISMATCH {m} {args} prefix={sharded.prefix} bounds={sharded.row_bounds}"))
        if rank == 0:
            queue.put(log)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_random_operation_sequences_on_three_ranks():
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fuzz_worker, args=(r, 3, port, queue)) for r in range(3)]
    for p in procs:
        p.start()
    log = queue.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bad = [entry for entry in log if entry[3] not in ("ok", "refused")]
    print(f"fuzz: {len(log)} checks, {sum(1 for e in log if e[3] == 'refused')} refused, ops {sorted({e[2] for e in log})}")
    assert not bad, bad[:5]
    assert sum(1 for entry in log if entry[3] == "ok") >= 60, len(log)
    assert {entry[2] for entry in log if entry[3] == "ok"} >= {"drillUp", "dice", "diceRange", "reorderDimensions", "removeDimension", "dice+rebalance"}
