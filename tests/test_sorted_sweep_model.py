"""A numpy MODEL of the bucket-sorted 2D histogram sweep (k_bin8c -> k_bucket_records ->
k_hist2d_records), kept next to the tests as an executable description of the algorithm: byte bins, the circular
anchor rule, one counting sort per anchor, 32-byte records in bucket order, per-bucket [bin][lane] accumulation in
64-bit fixed point.  It checks the DESIGN (every pair covered exactly once, grids == np.bincount bit for bit in fixed
point) on the CPU; the CUDA kernels are checked against np.bincount in tests/test_gpu_2d.py."""
import numpy as np
import pytest

from getdist_b200.parallel import _anchor_position


def circular_jobs(P):
    """anchor a owns the pairs with the next P/2 parameters (mod P); <= 32 partners per job"""
    jobs = []
    for a in range(P):
        partners = [(a + d) % P for d in range(1, P // 2 + 1) if (2 * d < P or a < (a + d) % P)]
        for k0 in range(0, len(partners), 32):
            jobs.append((a, partners[k0:k0 + 32]))
    return jobs


def sorted_sweep(bins, wq, jobs, G=256):
    """bins: (N, P) uint8, wq: (N,) uint64 fixed-point weights -> {(x, y): (G, G) uint64 grid [y][x]} for x < y"""
    grids = {}
    for a, partners in jobs:
        order = np.argsort(bins[:, a], kind="stable")        # counting sort by the anchor's bin (k_bucket_records)
        recs = bins[order][:, partners]                       # one record per row: the partners' bins in lane order
        w = wq[order]
        starts = np.searchsorted(bins[order, a], np.arange(G + 1))  # bucket boundaries (k_bucket_scan)
        for lane, b in enumerate(partners):
            x, y = min(a, b), max(a, b)
            grid = grids.setdefault((x, y), np.zeros((G, G), dtype=np.uint64))
            for c in range(G):                                # k_hist2d_records: rows of bucket c -> column / row c
                s, e = starts[c], starts[c + 1]
                if s == e:
                    continue
                col = np.zeros(G, dtype=np.uint64)            # the [bin][lane] shared-memory bins of this lane
                np.add.at(col, recs[s:e, lane], w[s:e])
                if a == x:
                    grid[:, c] += col                         # anchor is the x parameter: column c
                else:
                    grid[c, :] += col                         # anchor is the y parameter: row c
    return grids


@pytest.mark.parametrize("P", [2, 3, 8, 9])
def test_model_matches_bincount(P):
    rng = np.random.default_rng(P)
    N, G = 5000, 256
    bins = np.clip(rng.normal(128, 30, size=(N, P)), 0, 255).astype(np.uint8)
    wq = np.rint(rng.exponential(1.0, N) * 2.0**40).astype(np.uint64)
    jobs = circular_jobs(P)
    covered = sorted((min(a, b), max(a, b)) for a, ps in jobs for b in ps)
    assert covered == [(i, k) for i in range(P) for k in range(i + 1, P)]  # every pair exactly once
    grids = sorted_sweep(bins, wq, jobs, G)
    for (x, y), grid in grids.items():
        ref = np.zeros(G * G, dtype=np.uint64)
        np.add.at(ref, bins[:, x].astype(np.int64) + bins[:, y].astype(np.int64) * G, wq)
        assert np.array_equal(grid.reshape(-1), ref), (x, y)


@pytest.mark.parametrize("P", [4, 7, 64])
def test_partition_rule_is_the_library_rule(P):
    """getdist_b200.parallel splits the pair list by the same circular rule the library uses for a full triangle"""
    owner = {}
    for a, ps in circular_jobs(P):
        for b in ps:
            owner[(min(a, b), max(a, b))] = a
    for i in range(P):
        for k in range(i + 1, P):
            assert _anchor_position(i, k, P) == owner[(i, k)]
