"""Slab decomposition of axis 0 across ranks (one process per GPU).

Replaces the reference's in-process slab scheduler (solvers/iterator.cpp:68-84:
thread i owns rows [i*nX0/nT, (i+1)*nX0/nT) and is handed its slab of w plus one
recomputed ghost layer).  Here every rank owns a contiguous block of rows of u
and receives an N-row halo of u per side (N = order; boundaries.cpp:32-33 pads
by N), which is necessary and sufficient for the owned cells to see exactly the
stencils of the undivided domain.  The C++ driver implements the
same rules in Solver::exchange_halos / k_boundaries; this module is the host
mirror used to split/stitch arrays and by the gloo tests.
"""
import numpy as np

CLAMP, WRAP, HALO = 0, 1, 2   # GridParams.halo_lo / halo_hi in kernels.cuh


def slab_bounds(n0, rank, world):
    """Rows [start, stop) of axis 0 owned by `rank` (iterator.cpp:70-71)."""
    return rank * n0 // world, (rank + 1) * n0 // world


def halo_modes(rank, world, periodic):
    """(low, high) source of the N ghost rows of axis 0 for this rank."""
    if world == 1:
        m = WRAP if periodic else CLAMP
        return m, m
    lo = HALO if (rank > 0 or periodic) else CLAMP
    hi = HALO if (rank < world - 1 or periodic) else CLAMP
    return lo, hi


def neighbours(rank, world, periodic):
    """(low, high) neighbour ranks or None at a non-periodic domain edge."""
    lo = rank - 1 if rank > 0 else (world - 1 if periodic else None)
    hi = rank + 1 if rank < world - 1 else (0 if periodic else None)
    return lo, hi


def split(u, world):
    return [np.ascontiguousarray(u[slice(*slab_bounds(u.shape[0], r, world))])
            for r in range(world)]


def stitch(parts):
    return np.concatenate(parts, axis=0)


def padded_axis0(u_local, halo_lo, halo_hi, N, modes):
    """The axis-0 ghost rows of a slab as k_boundaries builds them."""
    lo_mode, hi_mode = modes
    n = u_local.shape[0]
    if lo_mode == HALO:
        lo = halo_lo
    elif lo_mode == WRAP:
        lo = u_local[n - N:]
    else:
        lo = np.repeat(u_local[:1], N, axis=0)
    if hi_mode == HALO:
        hi = halo_hi
    elif hi_mode == WRAP:
        hi = u_local[:N]
    else:
        hi = np.repeat(u_local[-1:], N, axis=0)
    return np.concatenate([lo, u_local, hi], axis=0)


def exchange_halos_torch(u_local, N, periodic):
    """Halo exchange over an initialised torch.distributed group (any backend):
    first N rows go to the low neighbour, last N rows to the high neighbour.
    Returns (halo_lo, halo_hi) as numpy arrays (None where there is no
    neighbour)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = neighbours(rank, world, periodic)
    first = torch.from_numpy(np.ascontiguousarray(u_local[:N]))
    last = torch.from_numpy(np.ascontiguousarray(u_local[-N:]))
    rlo = torch.empty_like(first)
    rhi = torch.empty_like(first)
    ops = []
    if lo is not None:
        ops += [dist.P2POp(dist.isend, first, lo), dist.P2POp(dist.irecv, rlo, lo)]
    if hi is not None:
        ops += [dist.P2POp(dist.isend, last, hi), dist.P2POp(dist.irecv, rhi, hi)]
    if world == 2 and lo is not None and hi is not None and lo == hi:
        # two ranks, periodic: both neighbours are the same peer; order the
        # messages by tag so low/high halos cannot be swapped
        reqs = [dist.isend(first, lo, tag=1), dist.isend(last, hi, tag=2),
                dist.irecv(rlo, lo, tag=2), dist.irecv(rhi, hi, tag=1)]
        for r in reqs:
            r.wait()
    elif ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return (rlo.numpy() if lo is not None else None, rhi.numpy() if hi is not None else None)
