"""N>1 host logic on CPU: a world_size-2 gloo group runs the density partition + all-gather with a stand-in
compute step and checks that every rank ends up with every density exactly once, in the caller's order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from getdist_b200.parallel import all_gather_grids, gather_order, partition_triangle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_density_1d(j, F):
    return torch.full((F,), float(j) + 0.5, dtype=torch.float64)


def _fake_density_2d(pr, G2):
    return torch.full((G2,), 1000.0 * pr[0] + pr[1], dtype=torch.float64)


def _worker(rank, world, port, P, F, G2, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = list(range(P))
    pairs = [(idx[i], idx[k]) for i in range(P) for k in range(i + 1, P)]
    my1d, my2d, max1d, per = partition_triangle(idx, pairs, rank, world)
    d1 = torch.zeros((max1d, F), dtype=torch.float64)
    d2 = torch.zeros((per, G2), dtype=torch.float64)
    for k, j in enumerate(my1d):
        d1[k] = _fake_density_1d(j, F)
    for k, pr in enumerate(my2d):
        d2[k] = _fake_density_2d(pr, G2)
    g1 = all_gather_grids(d1, world, dist)
    g2 = all_gather_grids(d2, world, dist)
    o1, o2 = gather_order(idx, pairs, world)
    ok = all(torch.equal(g1[o1[i]], _fake_density_1d(j, F)) for i, j in enumerate(idx))
    ok = ok and all(torch.equal(g2[o2[i]], _fake_density_2d(pr, G2)) for i, pr in enumerate(pairs))
    ret[rank] = bool(ok) and len(set(o1)) == len(idx) and len(set(o2)) == len(pairs)
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [5, 8])
def test_partition_and_allgather_world2(P):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), P, 16, 9, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def test_partition_covers_everything_once():
    for world in (1, 2, 3, 4, 8):
        for P in (1, 2, 7, 64):
            idx = list(range(P))
            pairs = [(i, k) for i in range(P) for k in range(i + 1, P)]
            seen1, seen2 = [], []
            for r in range(world):
                a, b, m1, per = partition_triangle(idx, pairs, r, world)
                assert len(a) <= m1 and len(b) <= per
                seen1 += a
                seen2 += b
            assert sorted(seen1) == idx and sorted(seen2) == sorted(pairs)


@pytest.mark.parametrize("P,world", [(64, 2), (64, 8), (9, 2), (33, 4)])
def test_split_pairs_keeps_anchors_together(P, world):
    """every pair appears once; all pairs of one anchor parameter land on one rank with a consistent hint; the
    per-rank counts differ by at most one anchor's worth of pairs"""
    from getdist_b200.parallel import _anchor_position, split_pairs

    idx = list(range(100, 100 + P))
    pairs = [(idx[i], idx[k]) for i in range(P) for k in range(i + 1, P)]
    lists, hints = split_pairs(idx, pairs, world)
    assert sorted(p for l in lists for p in l) == sorted(pairs)
    owner = {}
    for r in range(world):
        for (a, b), h in zip(lists[r], hints[r]):
            anchor = a if h == 1 else b
            assert _anchor_position(a - 100, b - 100, P) == anchor - 100
            assert owner.setdefault(anchor, r) == r
    counts = [len(l) for l in lists]
    blk = (P + world - 1) // world
    assert max(counts) - min(c for c in counts if c) <= blk * (P // 2 + 1)
    per_anchor = {}
    for r in range(world):
        for (a, b), h in zip(lists[r], hints[r]):
            per_anchor[a if h == 1 else b] = per_anchor.get(a if h == 1 else b, 0) + 1
    assert max(per_anchor.values()) <= P // 2 and min(per_anchor.values()) >= (P - 1) // 2


class _FakePar:
    def __init__(self):
        self._ranges_ready = False
        self.row = None


class _FakeMC:
    """stand-in with the three hooks exchange_param_ranges uses"""

    def __init__(self, P):
        import types

        self.paramNames = types.SimpleNamespace(names=[_FakePar() for _ in range(P)])
        self.calls = []
        self._ctx = types.SimpleNamespace(weighted_quantiles=self._wq)

    def _range_fracs(self):
        return np.linspace(0.05, 0.95, 11)

    def _wq(self, params, fr):
        self.calls.append(list(params))
        return np.array([[100.0 * j + f for f in fr] for j in params])

    def _finish_param(self, par, j, row):
        par.row = np.array(row)
        par._ranges_ready = True


def _range_worker(rank, world, port, P, ret):
    from getdist_b200.parallel import exchange_param_ranges

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mc = _FakeMC(P)
    exchange_param_ranges(mc, list(range(P)) + [0, 1], rank, world, dist, device="cpu")
    fr = mc._range_fracs()
    ok = all(p._ranges_ready and np.array_equal(p.row, 100.0 * j + fr) for j, p in enumerate(mc.paramNames.names))
    ok = ok and sum(len(c) for c in mc.calls) == len(range(rank, P, world))  # only this rank's share was selected here
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_exchange_param_ranges_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_range_worker, args=(world, _free_port(), 7, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]
