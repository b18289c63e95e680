"""The address tables of the fused rollup + exchange (olap_drill_up_rows, sharded.py): the numpy
construction against the plain loops it replaced, on random shard layouts (no GPU needed)."""
import numpy as np

from olap_in_memory_b200.sharded import _peer_row_tables, split_rows


def _by_loops(bases, out_bounds, me, K, inner, plane_v, plane_s, with_status):
    W, n = len(out_bounds) - 1, out_bounds[-1]
    rows_of = [out_bounds[r + 1] - out_bounds[r] for r in range(W)]
    owner = [next(r for r in range(W) if out_bounds[r] <= j < out_bounds[r + 1]) for j in range(n)]
    order = sorted(range(n), key=lambda j: ((owner[j] - me - 1) % W, j))
    position = [0] * n
    for q, j in enumerate(order):
        position[j] = q
    values, status = [0] * (K * n), [0] * (K * n)
    for k in range(K):
        for q, j in enumerate(order):
            r = owner[j]
            local = j - out_bounds[r]
            values[k * n + q] = bases[r] + k * plane_v + me * rows_of[r] * inner * 4 + local * inner * 4
            status[k * n + q] = bases[r] + K * plane_v + k * plane_s + me * rows_of[r] * inner + local * inner
    return position, values, status if with_status else None


def test_tables_match_the_loops():
    rng = np.random.default_rng(7)
    for _ in range(300):
        W = int(rng.integers(1, 9))
        n = int(rng.integers(1, 40))
        if rng.random() < 0.5:
            out_bounds = split_rows(n, W)
        else:  # uneven shards, some of them empty (after a dice of the sharded dimension)
            cuts = np.sort(rng.integers(0, n + 1, W - 1)).tolist()
            out_bounds = [0] + cuts + [n]
        me, K, inner = int(rng.integers(W)), int(rng.integers(1, 5)), int(rng.integers(1, 1000))
        r_max = max(b - a for a, b in zip(out_bounds, out_bounds[1:]))
        pad = lambda b: (b + 255) // 256 * 256
        plane_v, plane_s = pad(W * r_max * inner * 4), pad(W * r_max * inner)
        bases = [int(x) for x in rng.integers(1 << 40, 1 << 47, W)]  # device addresses are 47-bit
        with_status = bool(rng.integers(2))
        got = _peer_row_tables(bases, out_bounds, me, K, inner, plane_v, plane_s, with_status)
        want = _by_loops(bases, out_bounds, me, K, inner, plane_v, plane_s, with_status)
        assert got[0].dtype == np.int32 and got[0].tolist() == want[0]
        assert got[1].tolist() == want[1]
        assert (got[2] is None) == (want[2] is None) and (got[2] is None or got[2].tolist() == want[2])


def test_every_rank_starts_with_its_right_neighbour():
    W, inner, K = 4, 10, 1
    out_bounds = split_rows(8, W)
    bases = [1000 * (r + 1) << 20 for r in range(W)]
    for me in range(W):
        position, values, _ = _peer_row_tables(bases, out_bounds, me, K, inner, 4096, 1024, False)
        first_row = int(np.argmin(position))
        assert out_bounds[(me + 1) % W] == first_row  # the first table slot is the right neighbour's first row
        last_row = int(np.argmax(position))
        assert out_bounds[me] <= last_row < out_bounds[me + 1]  # ... and the walk ends with my own rows
        # all stores of rank `me` land in slot `me` of every receiver: no two senders share an address
        assert len(set(values.tolist())) == values.size
