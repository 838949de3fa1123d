"""Multi-GPU plumbing for the density batches: the path shards by independent units (densities), so each rank
computes a slice of the list and ONE all-gather of the result grids follows.  torch.distributed is
used for the collective only (NCCL on device tensors in production, gloo on CPU tensors in the tests)."""


def partition_triangle(idx, pairs, rank, world):
    """1D densities round-robin; 2D pairs in contiguous equal blocks (keeps the 8x8 histogram tiles dense).
    Returns (my1d, my2d, max1d, per2d): per2d / max1d are the padded per-rank counts used for the all-gather."""
    my1d = idx[rank::world]
    per = (len(pairs) + world - 1) // world
    my2d = pairs[rank * per: (rank + 1) * per]
    max1d = (len(idx) + world - 1) // world
    return my1d, my2d, max1d, per


def gather_order(idx, pairs, world):
    """Row index in the gathered (world*padded) tensors for every density, in the caller's order."""
    max1d = (len(idx) + world - 1) // world
    per = (len(pairs) + world - 1) // world
    rows1d = {}
    for r in range(world):
        for k, j in enumerate(idx[r::world]):
            rows1d[j] = r * max1d + k
    rows2d = {}
    for r in range(world):
        for k, pr in enumerate(pairs[r * per: (r + 1) * per]):
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
