"""The multi-GPU product path (getdist_b200.parallel.prefetch_triangle_group behind MCSamples(process_group=...)) on the
CPU: a world_size-2 gloo group, the library context replaced by a TEST DOUBLE whose "windows" are POSIX shared-memory
segments (the stand-in of CUDA IPC peer memory) and whose batch calls write a recognisable pattern into every rank's
window, as the library's peer stores do.  Checks the host logic of the path: window mapping through the rendezvous,
the gathered layouts (1D rows per rank, 2D grids in the caller's pair order), the result-record exchange, the quantile
exchange, and that every rank ends up with every density in the caller's order."""
import os
import socket
from multiprocessing import shared_memory

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from test_host_mirror_cpu import FakeContext


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def f1(j):
    return 0.25 + 0.001 * j


def f2(a, b):
    return 0.5 + 0.001 * a + 0.000001 * b


class WinContext(FakeContext):
    """numpy stand-in with windows: addresses are offsets into per-window shared-memory segments"""

    BASE = 1 << 44

    def __init__(self, device=0):
        super().__init__(device)
        self.own, self.peer = {}, {}
        self.rank, self.world = 0, 1

    def set_samples(self, X, w=None, chain_offsets=None, group=None):
        self.group_seen = group
        super().set_samples(X, w, chain_offsets)

    def peer_init(self, rank, world):
        self.rank, self.world = rank, world

    def window_export(self, window, nbytes):
        seg = shared_memory.SharedMemory(create=True, size=int(nbytes))
        self.own[window] = seg
        self.peer[window] = {}
        return self.BASE * (window + 1), seg.name.encode().ljust(64, b"\0")

    def window_import(self, window, peer, handle):
        self.peer[window][peer] = shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode())

    def _views(self, window):
        segs = [self.own[window]] + list(self.peer[window].values())
        return [np.ndarray((s.size // 8,), dtype=np.float64, buffer=s.buf) for s in segs]

    def stream_sync(self):
        pass

    def window_read(self, window, offset, out, sync=True):
        v = np.ndarray((self.own[window].size // 8,), dtype=np.float64, buffer=self.own[window].buf)
        out.reshape(-1)[:] = v[offset // 8: offset // 8 + out.size]
        return out

    def density1d_batch(self, specs, out=None, device_ptr=None, likes=False, stride=None, peers=False):
        from getdist_b200 import _abi

        res = []
        if out is not None:  # host output (the node's shared result buffer)
            views, off = [out.reshape(-1)], 0
        else:
            assert peers and device_ptr is not None
            views, off = self._views(_abi.GDK_WIN_G1), (device_ptr - self.BASE * (_abi.GDK_WIN_G1 + 1)) // 8
        for i, s in enumerate(specs):
            for v in views:
                v[off + i * stride: off + i * stride + s.fine_bins] = f1(s.param)
            r = _abi.Result1D()
            r.kde_h, r.status, r.winw = 0.1 * (s.param + 1), 0, s.param
            res.append(r)
        return None, res

    def density2d_batch(self, specs, out=None, device_ptr=None, likes=False, offsets=None, peers=False):
        from getdist_b200 import _abi

        if out is not None:
            views = [out]
        else:
            assert peers and device_ptr == self.BASE * (_abi.GDK_WIN_G2 + 1)
            views = self._views(_abi.GDK_WIN_G2)
        res = []
        for sp, off in zip(specs, offsets):
            G = int(sp["fine_bins"])
            for v in views:
                v[off: off + G * G] = f2(int(sp["px"]), int(sp["py"]))
            r = _abi.Result2D()
            r.hx, r.status, r.winw = 1.0 + int(sp["px"]), 0, int(sp["py"])
            for k in range(int(sp["n_contours"])):
                r.levels[k] = 100 * int(sp["px"]) + int(sp["py"]) + 0.1 * k
            res.append(r)
        return None, np.asarray(offsets), res

    def close(self):
        for seg in self.own.values():
            seg.close()
            seg.unlink()


def _worker(rank, world, port, P, root, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from getdist_b200 import MCSamples, _abi
    from getdist_b200.parallel import PeerGroup

    _abi.Context = WinContext
    _abi.host_register = lambda arr: None  # no CUDA here: the shared segment stays pageable
    rng = np.random.default_rng(5)
    N = 4000
    X = rng.standard_normal((N, P)) * np.arange(1, P + 1) + 10.0
    w = rng.exponential(1.0, N)
    pg = PeerGroup(dist, rank, world, device="cpu")
    r0, r1 = pg.row_range(10_000_000)
    mc = MCSamples(samples=X, weights=w, names=["p%d" % i for i in range(P)], sampler="uncorrelated",
                   settings={"fine_bins": 64, "fine_bins_2D": 16}, process_group=pg)
    d1, d2 = mc.prefetch_triangle(root=root)
    idx, pairs = mc.triangle_pairs()
    ok = pg.transport == "p2p" and mc._ctx.group_seen is pg
    if root is not None and rank != root:  # gather to one rank: the others hold nothing
        ret[rank] = (len(d1) == 0 and len(d2) == 0, [mc.paramNames.names[j].range_min for j in idx], (r0, r1))
        dist.barrier()
        mc._ctx.close()
        dist.destroy_process_group()
        return
    ok = ok and len(d1) == P and len(d2) == len(pairs)
    ok = ok and all(np.all(d.P == f1(j)) and d._gdk["winw"] == j for d, j in zip(d1, idx))
    ok = ok and all(np.all(d.P == f2(a, b)) and d._gdk["winw"] == b and d._gdk["hx"] == 1.0 + a
                    and d._gdk["levels"][1][0] == 100 * a + b for d, (a, b) in zip(d2, pairs))
    # the ranges came through the quantile exchange: identical on both ranks (checked by the parent)
    ret[rank] = (bool(ok), [mc.paramNames.names[j].range_min for j in idx], (r0, r1))
    dist.barrier()
    mc._ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("P,root", [(5, None), (8, None), (6, 0), (7, 1)])
def test_group_prefetch_world2(P, root):
    """root=None: every rank gets every density through the peer windows; root=r: the ranks copy their grids into the
    node's shared host buffer and only rank r wraps them"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), P, root, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == ret[1][1]
    # row blocks: whole statistics blocks, contiguous, covering [0, N)
    (a0, a1), (b0, b1) = ret[0][2], ret[1][2]
    assert a0 == 0 and a1 == b0 and b1 == 10_000_000 and a1 % 65536 == 0


def _worker_resident(rank, world, port, P, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from getdist_b200 import MCSamples, _abi
    from getdist_b200.parallel import PeerGroup, prefetch_triangle_group

    _abi.Context = WinContext
    rng = np.random.default_rng(7)
    N = 3000
    X = rng.standard_normal((N, P)) * np.arange(1, P + 1) - 3.0
    pg = PeerGroup(dist, rank, world, device="cpu")
    mc = MCSamples(samples=X, names=["p%d" % i for i in range(P)], sampler="uncorrelated",
                   settings={"fine_bins": 32, "fine_bins_2D": 8}, process_group=pg)
    idx, pairs = mc.triangle_pairs()
    ok = True
    for _ in range(2):  # the second call reuses the kept partition plan and the mapped windows
        mc.invalidate_density_caches()
        out = prefetch_triangle_group(mc, pg, idx, to_host=False)
        F = out["g1"]["stride"]
        g1 = mc._ctx.window_read(_abi.GDK_WIN_G1, 0, np.empty(world * ((P + world - 1) // world) * F))
        for j, row in zip(idx, out["g1"]["rows"]):
            ok = ok and bool(np.all(g1[row * F: (row + 1) * F] == f1(j)))
        offs, fb = out["g2"]["offsets"], out["g2"]["fine_bins"]
        g2 = mc._ctx.window_read(_abi.GDK_WIN_G2, 0, np.empty(int((fb * fb).sum())))
        for (a, b), o, G in zip(pairs, offs, fb):
            ok = ok and bool(np.all(g2[o: o + G * G] == f2(a, b)))
        # records of every rank's densities, in rank-major padded tables
        ok = ok and out["res1d"].shape[0] == world and out["res2d"].shape[:1] == (world,)
    ret[rank] = bool(ok)
    dist.barrier()
    mc._ctx.close()
    dist.destroy_process_group()


def test_group_resident_results_world2():
    """to_host=False (the bench's resident step): every rank's windows hold every grid at the rows / offsets the call
    reports"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_resident, args=(world, _free_port(), 7, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def test_group_prefetch_world3_odd_sizes():
    """three ranks, parameter counts that do not divide: padded 1D rows, anchor blocks of unequal size, gather to rank 1"""
    world = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 8, 1, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    assert ret[0][1] == ret[1][1] == ret[2][1]
