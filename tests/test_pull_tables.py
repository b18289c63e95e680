"""Child tables of the pull-model sharded rollup (olap_drill_up_pull, sharded._pull_tables): the numpy
construction against plain loops on random shard layouts and row maps (no GPU needed)."""
import numpy as np

from olap_in_memory_b200.sharded import _pull_tables, split_rows


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
