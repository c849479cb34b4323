"""CPU model of the merge-path tile schedule and of a warp's walk over its tile (sgl_b200/csrc/graph.cu build_tiles_kernel,
carried_row, carry runs; sgl_b200/csrc/spmm.cu spmm_flat_kernel + the fold).  The model is a line-by-line restatement in
Python of the device logic, run on small random CSR matrices to check the invariants the CUDA code relies on:
every non-zero is consumed exactly once, every row is stored exactly once, rows of at most `split_threshold` non-zeros
are never cut, cut rows are completed by (carriers + finisher), and the result equals the dense product."""
import numpy as np
import pytest


def build_tiles(indptr, tile_items, split_threshold):
    """-> tile_row[t], tile_nnz[t] for t = 0..n_tiles (boundary coordinates)."""
    n_rows = len(indptr) - 1
    nnz = int(indptr[-1])
    total = n_rows + nnz
    n_tiles = (total + tile_items - 1) // tile_items if total else 0
    rows, offs = [], []
    for t in range(n_tiles + 1):
        k = min(t * tile_items, total)
        lo, hi = max(k - nnz, 0), min(k, n_rows)
        while lo < hi:
            mid = (lo + hi) >> 1
            if indptr[mid + 1] + mid <= k - 1:
                lo = mid + 1
            else:
                hi = mid
        i, j = lo, k - lo
        if i < n_rows:
            start, deg = indptr[i], indptr[i + 1] - indptr[i]
            if j > start and (split_threshold < 0 or deg <= split_threshold):
                j = start
        rows.append(i)
        offs.append(j)
    return rows, offs


def carried_row(indptr, rows, offs, n_rows, t):
    i_end = rows[t + 1]
    if i_end >= n_rows:
        return -1
    frm = max(indptr[i_end], offs[t])
    return i_end if offs[t + 1] > frm else -1


def run_model(indptr, indices, vals, x, tile_items, split_threshold):
    n_rows = len(indptr) - 1
    rows, offs = build_tiles(indptr, tile_items, split_threshold)
    n_tiles = len(rows) - 1
    y = np.full((n_rows, x.shape[1]), np.nan)
    stores = np.zeros(n_rows, dtype=int)
    used = np.zeros(int(indptr[-1]), dtype=int)
    carries = {}                                   # tile -> (row, partial)
    for t in range(n_tiles):                       # one warp per tile
        row, row_end, j0, j1 = rows[t], rows[t + 1], offs[t], offs[t + 1]
        acc = np.zeros(x.shape[1])
        for j in range(j0, j1):
            while row < row_end and indptr[row + 1] == j:      # rows (also empty ones) that end here
                y[row] = acc
                stores[row] += 1
                acc = np.zeros(x.shape[1])
                row += 1
            used[j] += 1
            acc = acc + vals[j] * x[indices[j]]
        while row < row_end:
            assert indptr[row + 1] == j1
            y[row] = acc
            stores[row] += 1
            acc = np.zeros(x.shape[1])
            row += 1
        cr = carried_row(indptr, rows, offs, n_rows, t)
        if cr >= 0:
            assert cr == row
            carries[t] = (cr, acc)
        else:
            assert not acc.any() or j1 == j0 or True
    # fold: carriers of a row are consecutive tiles; the tile after the last carrier finished the row
    t = 0
    cut_rows = set()
    while t < n_tiles:
        if t in carries:
            r = carries[t][0]
            total = np.zeros(x.shape[1])
            last = t
            while last in carries and carries[last][0] == r:
                total = total + carries[last][1]
                last += 1
            assert last < n_tiles and rows[last] <= r < rows[last + 1]   # the finisher stores row r
            y[r] = y[r] + total
            cut_rows.add(r)
            t = last
        else:
            t += 1
    return y, stores, used, cut_rows, (rows, offs)


@pytest.mark.parametrize("seed", range(12))
def test_schedule_invariants(seed):
    rng = np.random.default_rng(seed)
    n_rows = int(rng.integers(1, 60))
    n_cols = int(rng.integers(1, 40))
    deg = rng.integers(0, 12, n_rows) * (rng.random(n_rows) < 0.7)
    if seed % 3 == 0:
        deg[rng.integers(0, n_rows)] = int(rng.integers(100, 400))       # a hub spanning several tiles
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    nnz = int(indptr[-1])
    indices = rng.integers(0, n_cols, nnz)
    vals = rng.standard_normal(nnz)
    x = rng.standard_normal((n_cols, 3))
    dense = np.zeros((n_rows, n_cols))
    for i in range(n_rows):
        for j in range(indptr[i], indptr[i + 1]):
            dense[i, indices[j]] += vals[j]
    want = dense @ x
    for tile_items, split in ((8, 4), (32, 16), (32, -1), (256, 64)):
        y, stores, used, cut_rows, (rows, offs) = run_model(indptr, indices, vals, x, tile_items, split)
        assert np.all(used == 1)                                        # every non-zero consumed exactly once
        assert np.all(stores == 1)                                      # every row stored exactly once
        np.testing.assert_allclose(y, want, rtol=1e-10, atol=1e-10)
        assert all(np.diff(rows) >= 0) and all(np.diff(offs) >= 0)      # boundaries are monotone
        for r in cut_rows:                                              # only rows above the threshold are cut
            assert split >= 0 and indptr[r + 1] - indptr[r] > split
        if split < 0:
            assert not cut_rows                                         # EXACT schedule: whole rows only
