"""Multi-GPU plumbing for the density batches (one process per GPU on one node).  The path shards by independent
units (densities), so every rank computes the densities of its block of anchor parameters and the grids are replicated
to all ranks -- there is no data-path collective in the usual sense:

  * `PeerGroup` maps every rank's result buffers (and, for the sharded upload, its sample store) into every other
    rank's address space with CUDA IPC handles exchanged once through torch.distributed.  The library then writes each
    finalised grid straight into all peers' gathered buffers from the kernel that normalises it (k_finalize2d /
    k_kde1d outputs over NVLink peer memory: the transfer overlaps the convolutions of the next group), and a rank's
    row block of a freshly uploaded sample matrix is transposed into every peer's column store by the same kernel that
    transposes it locally (upload 1/N of the rows per GPU over PCIe, NVLink for the rest).
  * Where IPC mapping is not available (no peer access) the group falls back to ONE NCCL all-gather of the result
    tensors after the batch -- loudly (`PeerGroup.transport == "nccl"`), never silently.

torch.distributed is used for the rendezvous, the tiny control collectives (quantile table, result structs, barrier)
and the fallback; gloo on CPU tensors in the tests."""
import numpy as np


def _anchor_position(pos_i, pos_k, P):
    """Position (in the parameter list) of the pair's anchor under the circular rule of the bucket-sorted 2D
    histogram sweep: parameter i owns the pairs with the next P/2 parameters (mod P)."""
    d = (pos_k - pos_i) % P
    return pos_i if (2 * d < P or (2 * d == P and pos_i < pos_k)) else pos_k


def split_pairs(idx, pairs, world):
    """Pairs grouped by rank so that every anchor's pairs stay on one rank (each rank then sweeps P/world full
    32-lane jobs instead of P partly filled ones).  Returns (lists, hints): hints[r][k] = 1 if the anchor of
    lists[r][k] is its x parameter, 2 if it is its y parameter (gdk_spec2d.anchor_hint)."""
    owner, hint = _pair_owners(idx, pairs, world)
    lists, hints = [], []
    for r in range(world):
        sel = np.nonzero(owner == r)[0].tolist()
        lists.append([pairs[n] for n in sel])
        hints.append(hint[sel].tolist())
    return lists, hints


def _pair_owners(idx, pairs, world):
    """(owner rank, anchor hint) of every pair, as arrays in the caller's pair order: _anchor_position and the block
    rule of split_pairs for all pairs at once (a P = 256 triangle has 32 640 of them)"""
    P = len(idx)
    if not len(pairs):
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    lut = {p: n for n, p in enumerate(idx)}
    pa = np.fromiter((lut[a] for a, _ in pairs), dtype=np.int64, count=len(pairs))
    pb = np.fromiter((lut[b] for _, b in pairs), dtype=np.int64, count=len(pairs))
    d = (pb - pa) % P
    first = (2 * d < P) | ((2 * d == P) & (pa < pb))
    anchor = np.where(first, pa, pb)
    blk = (P + world - 1) // world
    return np.minimum(anchor // blk, world - 1), np.where(first, 1, 2)


class TrianglePlan:
    """Everything about the partition of a triangle over the ranks that depends only on the parameter list and the world
    size: the pairs in the caller's order, their owners, each rank's list with its anchor hints, the padded per-rank
    counts and, for every density, its (rank, slot) in the gathered tables.  Built once per (parameter list, world) and
    kept (`triangle_plan`): a step of a 64-parameter triangle on 8 ranks is 20 ms, the Python loops over its 2016 pairs
    were 3 of them."""

    def __init__(self, idx, world, do_2d=True):
        self.idx, self.world = list(idx), int(world)
        n = len(self.idx)
        self.pairs = [(self.idx[i], self.idx[k]) for i in range(n) for k in range(i + 1, n)] if do_2d else []
        self.max1d = (n + world - 1) // world
        # 1D: round-robin; density n of the list is row (n % world) * max1d + n // world of the gathered rows
        pos = np.arange(n)
        self.rank1d, self.slot1d = pos % world, pos // world
        self.row1d = self.rank1d * self.max1d + self.slot1d
        self.jx = np.array([p[0] for p in self.pairs], dtype=np.int64)
        self.jy = np.array([p[1] for p in self.pairs], dtype=np.int64)
        owner, hint = _pair_owners(self.idx, self.pairs, world)
        self.rank2d = owner
        self.slot2d = np.zeros(len(self.pairs), dtype=np.int64)
        self.mine, self.hints, self.lists = [], [], []
        for r in range(world):
            sel = np.nonzero(owner == r)[0]
            self.slot2d[sel] = np.arange(sel.size)
            self.mine.append(sel)
            self.hints.append(hint[sel].astype(np.int32))
            self.lists.append([self.pairs[k] for k in sel.tolist()])
        self.per = max((m.size for m in self.mine), default=0)


_PLANS = {}


def triangle_plan(idx, world, do_2d=True):
    key = (tuple(idx), int(world), bool(do_2d))
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) >= 16:
            _PLANS.clear()
        plan = _PLANS[key] = TrianglePlan(idx, world, do_2d)
    return plan


def partition_triangle(idx, pairs, rank, world, with_hints=False):
    """1D densities round-robin; 2D pairs by anchor blocks (split_pairs).
    Returns (my1d, my2d, max1d, per2d): per2d / max1d are the padded per-rank counts used for the all-gather."""
    my1d = idx[rank::world]
    lists, hints = split_pairs(idx, pairs, world)
    per = max(len(l) for l in lists)
    max1d = (len(idx) + world - 1) // world
    if with_hints:
        return my1d, lists[rank], max1d, per, (hints[rank] if world > 1 else None)
    return my1d, lists[rank], max1d, per


def gather_order(idx, pairs, world):
    """Row index in the gathered (world*padded) tensors for every density, in the caller's order."""
    max1d = (len(idx) + world - 1) // world
    lists, _ = split_pairs(idx, pairs, world)
    per = max(len(l) for l in lists)
    rows1d = {}
    for r in range(world):
        for k, j in enumerate(idx[r::world]):
            rows1d[j] = r * max1d + k
    rows2d = {}
    for r in range(world):
        for k, pr in enumerate(lists[r]):
            rows2d[pr] = r * per + k
    return [rows1d[j] for j in idx], [rows2d[p] for p in pairs]


def all_gather_grids(local, world, dist=None):
    """local: (padded_count, grid_size) tensor on this rank -> (world*padded_count, grid_size) on every rank."""
    if world == 1:
        return local
    import torch

    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


def exchange_param_ranges(mc, indices, rank, world, dist=None, device="cuda"):
    """Exact weighted quantiles (the _initParam step, mcsamples.py:1427-1484) sharded over the ranks: every rank
    selects the order statistics of its share of the parameters on its own GPU, ONE small all-gather (11 doubles per
    parameter) hands every rank the full table, and the scalar range logic then runs everywhere.  The values are
    sample values selected by exact integer arithmetic, so every rank ends up with bit-identical ranges."""
    # every rank must pass the same set of parameters (and hold the same ready-state): the shares are dealt from
    # the sorted list
    todo = [j for j in sorted(set(indices)) if not mc.paramNames.names[j]._ranges_ready]
    if not todo:
        return
    if world == 1 or dist is None:
        mc._ensure_param_ranges(todo)
        return
    import torch

    fr = mc._range_fracs()
    mine = todo[rank::world]
    per = (len(todo) + world - 1) // world
    loc = np.zeros((per, fr.size))
    if mine:
        loc[: len(mine)] = mc._ctx.weighted_quantiles(mine, fr)
    send = torch.from_numpy(loc).to(device)
    recv = torch.empty((world * per, fr.size), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send)
    table = recv.cpu().numpy()  # the host planner needs the values: the one synchronisation of this step
    order = [j for r in range(world) for j in todo[r::world]]
    rows = [r * per + k for r in range(world) for k in range(len(todo[r::world]))]
    if hasattr(mc, "_finish_params"):
        mc._finish_params(order, table[rows])
    else:
        for j, row in zip(order, rows):
            mc._finish_param(mc.paramNames.names[j], j, table[row])


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) before it allocates page-locked buffers: with
    N ranks uploading at once, host buffers on the far socket halve the aggregate PCIe rate.  Returns the CPU list, or
    None where NVML or the affinity call is not available."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


class PeerGroup:
    """One process per GPU on one node: the rendezvous (torch.distributed: NCCL on GPUs, gloo in the CPU tests), this
    rank's id, and the bookkeeping of which library windows are mapped into the peers.

        dist.init_process_group("nccl", ...)
        pg = PeerGroup(dist, rank, world)
        mc = MCSamples(samples=X, weights=w, names=..., process_group=pg)   # sharded upload: 1/world of the rows per GPU
        d1, d2 = mc.prefetch_triangle()                                     # every rank ends up with every density

    transport: "p2p" = CUDA IPC windows (the product path); "nccl" = the loud fallback where peer mapping is not
    available (every rank uploads everything, ONE NCCL all-gather of the result tensors after the batch)."""

    def __init__(self, dist, rank, world, device="cuda", use_p2p=True):
        self.dist, self.rank, self.world, self.device = dist, int(rank), int(world), device
        self.p2p = bool(use_p2p) and self.world > 1
        self.transport = "single" if self.world == 1 else ("p2p" if self.p2p else "nccl")
        self.timings = {}

    # -- small control collectives ----------------------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def all_gather_array(self, arr):
        """host array per rank (same shape and dtype everywhere) -> (world,) + shape on every rank"""
        import torch

        arr = np.array(arr, copy=True)  # contiguous and writable (torch.from_numpy wants both)
        if self.world == 1:
            return arr[None]
        t = torch.from_numpy(arr.reshape(-1)).to(self.device)
        out = torch.empty((self.world * t.numel(),), dtype=t.dtype, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().reshape((self.world,) + arr.shape)

    def row_range(self, N):
        """rows this rank uploads itself: whole statistics blocks (GDK_ROW_BLOCK rows), dealt as evenly as they go"""
        from ._abi import GDK_ROW_BLOCK

        nblk = (N + GDK_ROW_BLOCK - 1) // GDK_ROW_BLOCK
        b0 = (self.rank * nblk) // self.world
        b1 = ((self.rank + 1) * nblk) // self.world
        return min(N, b0 * GDK_ROW_BLOCK), min(N, b1 * GDK_ROW_BLOCK)

    # -- peer windows -----------------------------------------------------------------------------------------
    def map_window(self, ctx, window, nbytes):
        """Window `window` of this rank's context with at least nbytes, mapped into every peer (every rank calls this
        with the same arguments).  Returns the local device address.  The windows only grow, and every rank sees the
        same sequence of sizes, so whether the handles must be exchanged again is decided without communication."""
        mapped = ctx.__dict__.setdefault("_peer_windows", {})  # lives and dies with the context
        have = mapped.get(window)
        if have is not None and have[1] >= nbytes:
            return have[0]
        ctx.peer_init(self.rank, self.world)
        addr, handle = ctx.window_export(window, nbytes)
        handles = self.all_gather_array(np.frombuffer(handle, dtype=np.uint8))
        ok = 1.0
        try:
            for r in range(self.world):
                if r != self.rank:
                    ctx.window_import(window, r, handles[r].tobytes())
        except Exception as e:  # no peer mapping on this box: every rank must learn it
            ok = 0.0
            self._why = str(e)
        if self.all_gather_array(np.array([ok])).min() < 1.0:
            self.p2p = False
            self.transport = "nccl (CUDA IPC peer mapping failed: %s)" % getattr(self, "_why", "on a peer")
            raise PeerMappingError(self.transport)
        mapped[window] = (addr, nbytes)
        return addr

    # -- shared host result buffer (gather to one rank) ---------------------------------------------------------
    def shared_results(self, nbytes, root):
        """A host buffer of at least nbytes that EVERY rank of the node maps (a /dev/shm file, unlinked once all have it
        open) and page-locks: each rank's batch call copies its grids straight into it over its own PCIe link, and the
        root rank reads them in place.  Collective.  The root reuses a segment once nothing refers to the arrays it
        handed out from it.  Returns a float64 array over the segment (the root's carries the reference that keeps the
        segment busy)."""
        import mmap
        import os
        import uuid
        import weakref

        from . import _abi

        segs = self.__dict__.setdefault("_segs", {})
        msg = np.zeros(72, dtype=np.uint8)
        if self.rank == root:
            seg = None
            for sg in segs.values():
                if sg["owner"] and not sg["busy"] and sg["size"] >= nbytes and (seg is None or sg["size"] < seg["size"]):
                    seg = sg
            if seg is None:
                size = ((int(nbytes * 1.05) + (2 << 20) - 1) // (2 << 20)) * (2 << 20)
                name = "gdk_%s" % uuid.uuid4().hex
                fd = os.open("/dev/shm/" + name, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, size)
                seg = dict(name=name, size=size, fd=fd, owner=True, busy=False, mm=None)
            nm = seg["name"].encode()
            msg[: len(nm)] = np.frombuffer(nm, dtype=np.uint8)
            msg[64:72] = np.frombuffer(np.int64(seg["size"]).tobytes(), dtype=np.uint8)
        row = self.all_gather_array(msg)[root]
        name = bytes(row[:64]).rstrip(b"\0").decode()
        size = int(np.frombuffer(bytes(row[64:72]), dtype=np.int64)[0])
        sg = segs.get(name)
        fresh = sg is None or sg["mm"] is None
        if fresh:
            if self.rank == root:
                sg = seg
            else:
                sg = dict(name=name, size=size, fd=os.open("/dev/shm/" + name, os.O_RDWR), owner=False, busy=False, mm=None)
            sg["mm"] = mmap.mmap(sg["fd"], size)
            os.close(sg["fd"])
            sg["arr"] = np.frombuffer(sg["mm"], dtype=np.float64)
            ok = 1.0
            try:
                _abi.host_register(sg["arr"])
            except Exception:  # page-locking refused (limits on locked memory): every rank must learn it
                ok = 0.0
            segs[name] = sg
            allok = self.all_gather_array(np.array([ok]))  # also the rendezvous after which the name can go
            if self.rank == root:
                os.unlink("/dev/shm/" + name)
            if allok.min() < 1.0:
                if ok:
                    _abi.host_unregister(sg["arr"])
                sg["arr"] = None
                sg["mm"].close()
                sg["mm"] = None
                self.host_gather = False  # from now on: gather through the root's device window
                return None
        if self.rank != root:
            return sg["arr"]
        sg["busy"] = True
        base = np.frombuffer(sg["mm"], dtype=np.float64)  # a fresh base object per call: its death frees the segment
        weakref.finalize(base, sg.__setitem__, "busy", False)
        return base

    def probe(self, ctx):
        """decide the transport once: try to map a small result window into the peers"""
        if self.world == 1 or not self.p2p or getattr(self, "_probed", False):
            return self.p2p
        self._probed = True
        try:
            from ._abi import GDK_WIN_G1

            self.map_window(ctx, GDK_WIN_G1, 1 << 16)
        except PeerMappingError:
            import logging

            logging.getLogger("getdist_b200").warning("multi-GPU transport: %s", self.transport)
        return self.p2p


class PeerMappingError(RuntimeError):
    pass


def _res1d_table(res):
    return np.array([[r.kde_h, r.h_raw, r.smooth_1D, r.winw, r.status, r.n_feval] for r in res], dtype=np.float64).reshape(-1, 6)


def _res2d_table(res):
    import ctypes

    from ._abi import RES2D_DTYPE

    if isinstance(res, ctypes.Array):  # the library's record array: column-wise, no Python loop
        a = np.frombuffer(res, dtype=RES2D_DTYPE)
        out = np.empty((a.size, 13))
        for i, k in enumerate(("hx", "hy", "c", "rx", "ry", "t_star", "winw", "status", "n_brent")):
            out[:, i] = a[k]
        out[:, 9:13] = a["levels"]
        return out
    return np.array([[r.hx, r.hy, r.c, r.rx, r.ry, r.t_star, r.winw, r.status, r.n_brent] + list(r.levels) for r in res],
                    dtype=np.float64).reshape(-1, 13)


def _res1d_from(row):
    from ._abi import Result1D

    r = Result1D()
    r.kde_h, r.h_raw, r.smooth_1D = row[0], row[1], row[2]
    r.winw, r.status, r.n_feval = int(row[3]), int(row[4]), int(row[5])
    return r


def _res2d_from(row):
    from ._abi import Result2D

    r = Result2D()
    r.hx, r.hy, r.c, r.rx, r.ry, r.t_star = row[:6]
    r.winw, r.status, r.n_brent = int(row[6]), int(row[7]), int(row[8])
    for k in range(4):
        r.levels[k] = row[9 + k]
    return r


def prefetch_triangle_group(mc, pg, idx, do_1d=True, do_2d=True, to_host=True, root=None):
    """MCSamples.prefetch_triangle over a PeerGroup: the densities of the triangle are partitioned across the ranks
    (1D round-robin, 2D by anchor blocks), every rank computes its share from its resident copy of the samples and the
    library stores each finished grid into the gathered windows of ALL ranks over NVLink while the next group is
    convolved; the per-density result records of both batches travel in ONE small all-gather, which is also the closing
    rendezvous (a rank contributes after its library calls have returned).  After it every rank
    holds every density.  to_host=False leaves the grids in the device windows (returns their addresses and layout).
    root=r gathers to rank r only (the process that plots): the grids are stored into its window alone and the other
    ranks return empty lists -- N host copies of a gigabyte of grids share one host memory system."""
    import time

    from . import _abi

    t0 = time.perf_counter()
    rank, world = pg.rank, pg.world
    if not pg.probe(mc._ctx):
        return _prefetch_triangle_nccl(mc, pg, idx, do_1d, do_2d, to_host)
    if hasattr(mc._ctx, "peer_targets"):
        mc._ctx.peer_targets(0xFFFFFFFF if root is None else (1 << int(root)))
    mine_host = to_host and (root is None or int(root) == rank)
    host_gather = to_host and root is not None and hasattr(pg, "shared_results") and getattr(pg, "host_gather", True)
    exchange_param_ranges(mc, idx, rank, world, pg.dist, pg.device)
    if mc.smooth_scale_1D <= 0 or mc.smooth_scale_2D < 0:
        mc._ensure_neff(idx)
    plan = triangle_plan(idx, world, do_2d)  # the partition: depends on the parameter list and the world size only
    pairs, max1d, per = plan.pairs, plan.max1d, plan.per
    my1d = idx[rank::world]
    out = {"transport": pg.transport}
    t1 = time.perf_counter()
    # ---- 1D: window rows [r * max1d, (r + 1) * max1d) belong to rank r
    d1 = []
    specs_all = None
    if do_1d:
        F = int(mc.fine_bins)
        # the wrapping rank needs every spec (Density1D axes and ranges), the others only their own
        specs_all = [mc._spec_1d(j, {}) for j in idx] if mine_host else None
        my_specs = specs_all[rank::world] if mine_host else [mc._spec_1d(j, {}) for j in my1d]
        tab = np.zeros((max1d, 6))
        n1 = world * max1d * F
        shared = None
        if host_gather:
            # one host buffer for the node: [1D rows per rank | 2D grids in the caller's pair order]
            fbq = mc._fine_bins_2d_all(pairs, plan.jx, plan.jy) if pairs else np.zeros(0, dtype=np.int64)
            shared = pg.shared_results((n1 + int((fbq * fbq).sum())) * 8, int(root))
            if shared is None:
                host_gather = False
        if host_gather:
            if my1d:
                _, res = mc._ctx.density1d_batch(my_specs, out=shared[rank * max1d * F: (rank * max1d + len(my1d)) * F].reshape(len(my1d), F),
                                                 stride=F)
                tab[: len(my1d)] = _res1d_table(res)
        else:
            base1 = pg.map_window(mc._ctx, _abi.GDK_WIN_G1, n1 * 8)
            if my1d:
                _, res = mc._ctx.density1d_batch(my_specs, device_ptr=base1 + rank * max1d * F * 8, stride=F, peers=True)
                tab[: len(my1d)] = _res1d_table(res)
        tab1 = tab
        if not pairs:
            out["res1d"] = pg.all_gather_array(tab1)
    t2 = time.perf_counter()
    # ---- 2D: the gathered buffer holds every pair's grid in the caller's pair order
    wrapped = None
    if pairs:
        conts = [float(c) for c in list(mc.contours[:4])]
        fb = mc._fine_bins_2d_all(pairs, plan.jx, plan.jy)
        offs = np.zeros(len(pairs), dtype=np.int64)
        offs[1:] = np.cumsum(fb * fb)[:-1]
        total = int((fb * fb).sum())
        my2d, mine, hints = plan.lists[rank], plan.mine[rank], plan.hints[rank]
        tab = np.zeros((per, 14))
        n1 = world * max1d * int(mc.fine_bins) if do_1d else 0
        if host_gather and shared is None:
            shared = pg.shared_results(total * 8, int(root))
            if shared is None:
                host_gather = False
        if host_gather and shared.size < n1 + total:
            raise RuntimeError("shared result buffer smaller than the gathered layout")
        if not host_gather:
            base2 = pg.map_window(mc._ctx, _abi.GDK_WIN_G2, total * 8)

        def full_specs():
            sp_all = mc._specs_2d_batch(pairs, {})
            assert np.array_equal(sp_all["fine_bins"].astype(np.int64), fb)
            return sp_all

        def wrap_all():  # host side of the wrapping rank: runs while this rank's library call is in flight
            return mc._wrap_2d(pairs, full_specs(), shared[n1: n1 + total], offs, conts)

        if len(mine):
            sp = mc._specs_2d_batch(my2d, {})  # only this rank's pairs on the critical path
            sp["n_contours"] = len(conts)
            for k, c in enumerate(conts):
                sp["contours"][:, k] = c
            sp["anchor_hint"] = hints
            if host_gather:  # every rank copies its grids into the node's shared host buffer over its own PCIe link
                call = lambda: mc._ctx.density2d_batch(sp, out=shared[n1: n1 + total], offsets=offs[mine])  # noqa: E731
                if mine_host:
                    from .mcsamples import _overlapped

                    (_, _, res), wrapped = _overlapped(call, wrap_all)
                else:
                    _, _, res = call()
            else:
                _, _, res = mc._ctx.density2d_batch(sp, device_ptr=base2, offsets=offs[mine], peers=True)
            tab[: len(mine), :13] = _res2d_table(res)
            tab[: len(mine), 13] = sp["bw_mode"]
        elif host_gather and mine_host:
            wrapped = wrap_all()
        # ONE exchange for the result records of both batches.  It is also the closing rendezvous: a rank contributes after
        # its library calls have returned (streams synchronised, its stores into every window and into the shared host
        # segment complete), so once the gathered table is here every rank's grids are where the readers expect them
        if do_1d:
            both = pg.all_gather_array(np.concatenate([tab1.ravel(), tab.ravel()]))
            out["res1d"] = both[:, : tab1.size].reshape((world,) + tab1.shape)
            out["res2d"] = both[:, tab1.size:].reshape((world,) + tab.shape)
        else:
            out["res2d"] = pg.all_gather_array(tab)
    t3 = time.perf_counter()
    t4 = t3
    d2 = []
    t_d2h = 0.0
    if do_1d:
        if mine_host:
            rec1 = out["res1d"][plan.rank1d, plan.slot1d]  # (len(idx), 6) in the caller's order
            if host_gather:
                P1 = shared[:n1].reshape(world * max1d, F)
            else:
                P1 = mc._ctx.window_read(_abi.GDK_WIN_G1, 0, _abi.result_buffer(world * max1d * F).reshape(world * max1d, F))
            d1 = mc._finish_1d(idx, specs_all, [P1[r] for r in plan.row1d.tolist()], [_res1d_from(row) for row in rec1])
        elif not to_host:
            out["g1"] = dict(address=base1, stride=F, rows=plan.row1d.tolist())
    if pairs:
        if mine_host:
            tr = time.perf_counter()
            tab2 = out["res2d"][plan.rank2d, plan.slot2d]  # (pairs, 14) in the caller's order -> result columns
            names2 = ("hx", "hy", "c", "rx", "ry", "t_star", "winw", "status", "n_brent")
            rcol = {k: (tab2[:, i].astype(np.int64).tolist() if k in ("winw", "status", "n_brent") else tab2[:, i].tolist())
                    for i, k in enumerate(names2)}
            rcol["levels"] = tab2[:, 9:13].tolist()
            if host_gather:  # the grids are already on this host (every rank copied its share) and wrapped
                d2, cols = wrapped
                mc._records_2d(pairs, cols, rcol)
            else:
                # the copy runs while the host wraps the grids (views of the buffer: nothing reads it before the sync below)
                buf = mc._ctx.window_read(_abi.GDK_WIN_G2, 0, _abi.result_buffer(total), sync=False)
                d2 = mc._finish_2d(pairs, full_specs(), buf, offs, rcol, conts)
                mc._ctx.stream_sync()
            t_d2h = time.perf_counter() - tr
        elif not to_host:
            out["g2"] = dict(address=base2, offsets=offs, fine_bins=fb)
    t5 = time.perf_counter()
    pg.timings = dict(ranges_ms=(t1 - t0) * 1e3, d1_ms=(t2 - t1) * 1e3, d2_ms=(t3 - t2) * 1e3, barrier_ms=(t4 - t3) * 1e3,
                      read_ms=(t5 - t4) * 1e3, d2h_2d_ms=t_d2h * 1e3)
    if to_host:
        return d1, d2
    return out


def _prefetch_triangle_nccl(mc, pg, idx, do_1d, do_2d, to_host):
    """fallback transport: every rank computes its share into local tensors, ONE all-gather per result tensor"""
    import torch

    rank, world = pg.rank, pg.world
    exchange_param_ranges(mc, idx, rank, world, pg.dist, pg.device)
    pairs = [(idx[i], idx[k]) for i in range(len(idx)) for k in range(i + 1, len(idx))] if do_2d else []
    my1d, my2d, max1d, per, hints = partition_triangle(idx, pairs, rank, world, with_hints=True) if pairs else (
        idx[rank::world], [], (len(idx) + world - 1) // world, 0, None)
    rows1d, rows2d = gather_order(idx, pairs, world) if pairs else ([(n % world) * max1d + n // world for n in range(len(idx))], [])
    d1, d2 = [], []
    if do_1d:
        F = int(mc.fine_bins)
        loc = torch.zeros((max1d, F), dtype=torch.float64, device=pg.device)
        tab = np.zeros((max1d, 6))
        specs_all = [mc._spec_1d(j, {}) for j in idx]
        pos = {j: n for n, j in enumerate(idx)}
        if my1d:
            _, res = mc._ctx.density1d_batch([specs_all[pos[j]] for j in my1d], device_ptr=loc.data_ptr(), stride=F)
            tab[: len(my1d)] = _res1d_table(res)
        g1 = all_gather_grids(loc, world, pg.dist).cpu().numpy()
        t1 = pg.all_gather_array(tab).reshape(world * max1d, 6)
        d1 = mc._finish_1d(idx, specs_all, [g1[r] for r in rows1d], [_res1d_from(t1[r]) for r in rows1d])
    if pairs:
        specs = mc._specs_2d_batch(pairs, {})
        conts = [float(c) for c in list(mc.contours[:4])]
        specs["n_contours"] = len(conts)
        for k, c in enumerate(conts):
            specs["contours"][:, k] = c
        G2 = int(specs["fine_bins"].astype(np.int64).max()) ** 2  # padded slots of the largest grid
        where = {pr: n for n, pr in enumerate(pairs)}
        mine = np.array([where[pr] for pr in my2d], dtype=np.int64)
        loc = torch.zeros((max(per, 1), G2), dtype=torch.float64, device=pg.device)
        tab = np.zeros((per, 13))
        if len(mine):
            sp = np.ascontiguousarray(specs[mine])
            sp["anchor_hint"] = np.asarray(hints, dtype=np.int32)
            _, _, res = mc._ctx.density2d_batch(sp, device_ptr=loc.data_ptr(), offsets=np.arange(len(mine), dtype=np.int64) * G2)
            tab[: len(mine)] = _res2d_table(res)
        g2 = all_gather_grids(loc, world, pg.dist).cpu().numpy().reshape(-1)
        t2 = pg.all_gather_array(tab).reshape(world * per, 13)
        d2 = mc._finish_2d(pairs, specs, g2, np.array(rows2d, dtype=np.int64) * G2, [_res2d_from(t2[r]) for r in rows2d], conts)
    return d1, d2
