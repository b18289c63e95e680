"""Child tables of the pull-model sharded rollup (olap_drill_up_pull, sharded._pull_tables): the numpy
construction against plain loops on random shard layouts and row maps (no GPU needed)."""
import numpy as np

from olap_in_memory_b200.sharded import _exchange_costs, _pull2_tables, _pull_tables, split_rows


def _by_loops(full_map, in_bounds, j0, j1):
    W = len(in_bounds) - 1
    row_start, rank, row = [0], [], []
    for j in range(j0, j1):
        for i, parent in enumerate(full_map):  # ascending global input row
            if parent == j:
                r = next(q for q in range(W) if in_bounds[q] <= i < in_bounds[q + 1])
                rank.append(r)
                row.append(i - in_bounds[r])
        row_start.append(len(rank))
    return row_start, rank, row


def _random_bounds(rng, n, W):
    if rng.random() < 0.5:
        return split_rows(n, W)
    return [0] + np.sort(rng.integers(0, n + 1, W - 1)).tolist() + [n]  # uneven, some shards empty


def test_tables_match_the_loops():
    rng = np.random.default_rng(11)
    for _ in range(300):
        W = int(rng.integers(1, 9))
        n_in, n_out = int(rng.integers(1, 60)), int(rng.integers(1, 25))
        full_map = rng.integers(0, n_out, n_in)
        in_bounds, out_bounds = _random_bounds(rng, n_in, W), _random_bounds(rng, n_out, W)
        seen = 0
        for me in range(W):
            j0, j1 = out_bounds[me], out_bounds[me + 1]
            got = _pull_tables(full_map, in_bounds, j0, j1)
            want = _by_loops(full_map.tolist(), in_bounds, j0, j1)
            assert got[0].dtype == np.int32 and got[1].dtype == np.int32 and got[2].dtype == np.int64
            assert got[0].tolist() == want[0] and got[1].tolist() == want[1] and got[2].tolist() == want[2]
            seen += len(want[1])
        assert seen == n_in  # every input row is the child of exactly one output row of exactly one rank


def test_children_are_walked_in_global_row_order():
    # prefix (d0=3, d1=2) rolled up on d0 -> output rows = d1; children of row j are d0*2 + j, ascending d0
    full_map = np.array([0, 1, 0, 1, 0, 1])
    row_start, rank, row = _pull_tables(full_map, [0, 2, 4, 6], 0, 2)
    assert row_start.tolist() == [0, 3, 6]
    assert rank.tolist() == [0, 1, 2, 0, 1, 2] and row.tolist() == [0, 0, 0, 1, 1, 1]


def test_exchange_costs_and_two_phase_tables_match_the_loops():
    rng = np.random.default_rng(3)
    for _ in range(200):
        W = int(rng.integers(1, 9))
        n_in, n_out = int(rng.integers(1, 60)), int(rng.integers(1, 25))
        full_map = rng.integers(0, n_out, n_in)
        in_bounds, out_bounds = _random_bounds(rng, n_in, W), _random_bounds(rng, n_out, W)
        direct, partial, partial_rows, local_rows, touched = _exchange_costs(full_map, in_bounds, out_bounds)
        owner_in = [next(q for q in range(W) if in_bounds[q] <= i < in_bounds[q + 1]) for i in range(n_in)]
        owner_out = [next(q for q in range(W) if out_bounds[q] <= j < out_bounds[q + 1]) for j in range(n_out)]
        want_touched = [sorted({int(full_map[i]) for i in range(n_in) if owner_in[i] == s}) for s in range(W)]
        assert [t.tolist() for t in touched] == want_touched
        assert direct == max(sum(1 for i in range(n_in) if owner_out[full_map[i]] == r and owner_in[i] != r) for r in range(W))
        assert partial == max(sum(1 for s in range(W) if s != r for j in want_touched[s] if owner_out[j] == r) for r in range(W))
        assert partial_rows == max(len(t) for t in want_touched) and local_rows == max(b - a for a, b in zip(in_bounds, in_bounds[1:]))
        for me in range(W):
            j0, j1 = out_bounds[me], out_bounds[me + 1]
            row_start, ranks, rows = _pull2_tables(touched, j0, j1)
            want_rs, want_rank, want_row = [0], [], []
            for j in range(j0, j1):
                for s in range(W):
                    if j in want_touched[s]:
                        want_rank.append(s)
                        want_row.append(want_touched[s].index(j))
                want_rs.append(len(want_rank))
            assert row_start.tolist() == want_rs and ranks.tolist() == want_rank and rows.tolist() == want_row


def test_the_plan_prefers_partials_when_ranks_hold_many_children():
    # univac cube, rows (d0, d1), dim0 -> all: 2 ranks hold 5 d0 values each -> partials are 5x smaller than the children
    full_map = np.arange(100) % 10
    direct, partial, *_ = _exchange_costs(full_map, split_rows(100, 2), split_rows(10, 2))
    assert (direct, partial) == (25, 5)
    # 8 ranks, deepened to (d0, d1, d2): about one d0 value per rank -> pulling the children themselves moves less
    full_map = np.arange(1000) % 100
    direct, partial, *_ = _exchange_costs(full_map, [b * 10 for b in split_rows(100, 8)], split_rows(100, 8))
    assert (direct, partial) == (117, 91)
    assert direct * 3 < partial * 4  # 3 measure planes against 4 partial planes (sum, average as sum + count, highest)
