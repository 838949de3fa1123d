"""Multi-GPU plumbing for the density batches: the path shards by independent units (densities), so each rank
computes a slice of the list and ONE all-gather of the result grids follows.  torch.distributed is
used for the collective only (NCCL on device tensors in production, gloo on CPU tensors in the tests)."""


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
    import numpy as np

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
    table = recv.cpu().numpy()
    for r in range(world):
        for k, j in enumerate(todo[r::world]):
            mc._finish_param(mc.paramNames.names[j], j, table[r * per + k])
