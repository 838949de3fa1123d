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
    P = len(idx)
    pos = {p: n for n, p in enumerate(idx)}
    blk = (P + world - 1) // world
    lists = [[] for _ in range(world)]
    hints = [[] for _ in range(world)]
    for (a, b) in pairs:
        ap = _anchor_position(pos[a], pos[b], P)
        r = min(ap // blk, world - 1)
        lists[r].append((a, b))
        hints[r].append(1 if ap == pos[a] else 2)
    return lists, hints


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
    for r in range(world):
        for k, j in enumerate(todo[r::world]):
            mc._finish_param(mc.paramNames.names[j], j, table[r * per + k])


class PeerGroup:
    """One process per GPU on one node.  Holds the rendezvous (torch.distributed), this rank's id, and -- per
    MCSamples object attached to it -- the gathered result buffers of the library, mapped into every peer.

        pg = PeerGroup(dist, rank, world)                       # after dist.init_process_group("nccl")
        d1, d2 = mc.prefetch_triangle(process_group=pg)         # every rank ends up with every density
    """

    SLOT_G1, SLOT_G2, SLOT_X, SLOT_W, SLOT_WQ = 0, 1, 2, 3, 4

    def __init__(self, dist, rank, world, device="cuda", use_p2p=True):
        self.dist, self.rank, self.world, self.device = dist, rank, world, device
        self.use_p2p = use_p2p and world > 1
        self.transport = "single" if world == 1 else ("p2p" if self.use_p2p else "nccl")
        self._bufs = {}  # (id(ctx), slot) -> (ptr, bytes)
        self.last = {}

    # -- small control collectives ----------------------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def all_gather_bytes(self, payload, nbytes):
        """fixed-size byte strings from every rank (IPC handles)"""
        import torch

        if self.world == 1:
            return [payload]
        t = torch.frombuffer(bytearray(payload.ljust(nbytes, b"\0")), dtype=torch.uint8).to(self.device)
        out = torch.empty(self.world * nbytes, dtype=torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        raw = out.cpu().numpy().tobytes()
        return [raw[r * nbytes:(r + 1) * nbytes] for r in range(self.world)]

    def all_gather_array(self, arr):
        """(n, k) float64 host array per rank (same shape everywhere) -> (world, n, k)"""
        import torch

        if self.world == 1:
            return arr[None]
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()

    # -- peer windows -----------------------------------------------------------------------------------------
    def window(self, ctx, slot, nbytes):
        """Device buffer `slot` of this context with at least nbytes, mapped into every peer (same size on every rank:
        all ranks call this with the same arguments).  Returns the local device pointer."""
        key = (id(ctx), slot)
        have = self._bufs.get(key)
        if have is not None and have[1] >= nbytes:
            return have[0]
        ptr, handle = ctx.peer_alloc(slot, nbytes)
        if self.use_p2p:
            handles = self.all_gather_bytes(handle, 64)
            try:
                for r, h in enumerate(handles):
                    if r != self.rank:
                        ctx.peer_open(slot, r, h)
            except Exception as e:  # no peer access on this box: say so and use the collective
                self.use_p2p = False
                self.transport = "nccl (CUDA IPC mapping failed: %s)" % e
            ok = self.all_gather_array(np.array([[1.0 if self.use_p2p else 0.0]]))
            if ok.min() < 1.0 and self.use_p2p:
                self.use_p2p = False
                self.transport = "nccl (CUDA IPC mapping failed on a peer)"
        self._bufs[key] = (ptr, nbytes)
        return ptr

    def tensor(self, ptr, shape):
        """torch view (float64) of a library-owned device buffer"""
        import torch

        n = int(np.prod(shape))

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}

        return torch.as_tensor(_Arr(), device=self.device).view(*shape)
