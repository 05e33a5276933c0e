"""One-cell halo exchange of the LATERALFLOW pass-1 planes (KCELL, HEAD) between neighbouring tiles.

The reference's MPI build never exchanges this halo (its mpp_land_com* routines are dead code and every rank
passes ids=its), so its opt_run=5 answer depends on the rank count.  Here the tiles reproduce the
SEQUENTIAL single-domain result: between noahmp_b200_wtable_begin and _end every rank sends its edge cells to the
up-to-8 neighbours of the mpp_land process grid.  Two phases (left/right columns, then up/down rows including the
freshly received ring columns) deliver the corners without diagonal messages.  Works on any torch.distributed
backend: NCCL on the device planes (NVLink), gloo on CPU tensors in the tests.
"""
import torch
import torch.distributed as dist

from .driver import proc_grid


def neighbours(rank, world):
    npx, npy = proc_grid(world)
    ipx, ipy = rank % npx, rank // npx
    left = rank - 1 if ipx > 0 else None
    right = rank + 1 if ipx < npx - 1 else None
    down = rank - npx if ipy > 0 else None
    up = rank + npx if ipy < npy - 1 else None
    return left, right, down, up


def allreduce_budget(sums):
    """Host-side form of the global water / energy budget for a driver whose ranks talk through torch.distributed
    (any backend) instead of the library's own NCCL communicator: sums = the dict NoahMP.budget_read() returns for this
    tile; returns the dict summed over all ranks ('steps' is the common step count, not a sum)."""
    names = list(sums)
    v = torch.tensor([float(sums[n]) for n in names], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            v = v.cuda()
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        v = v.cpu()
        out = dict(zip(names, v.tolist()))
        if "steps" in out:
            out["steps"] /= dist.get_world_size()
        return out
    return dict(zip(names, v.tolist()))


def _exchange(pairs):
    """pairs: list of (peer, send_tensor, recv_tensor). Ordered so that lower ranks send first (deadlock-free on
    backends without batched p2p); uses batch_isend_irecv where available."""
    ops = []
    for peer, snd, rcv in pairs:
        ops.append(dist.P2POp(dist.isend, snd, peer))
        ops.append(dist.P2POp(dist.irecv, rcv, peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_halo(kcell, head, rank=None, world=None):
    """kcell, head: torch tensors of shape (nj+2, ni+2) (device planes of NoahMP.wtable_halo(), or CPU tensors)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return
    left, right, down, up = neighbours(rank, world)
    planes = (kcell, head)
    # phase 1: columns (interior rows only)
    pairs, post = [], []
    for peer, s_col, r_col in ((left, 1, 0), (right, -2, -1)):
        if peer is None:
            continue
        snd = torch.stack([p[1:-1, s_col] for p in planes]).contiguous()
        rcv = torch.empty_like(snd)
        pairs.append((peer, snd, rcv))
        post.append((rcv, r_col))
    _exchange(pairs)
    for rcv, r_col in post:
        for k, p in enumerate(planes):
            p[1:-1, r_col] = rcv[k]
    # phase 2: rows, ring columns included (carries the corners)
    pairs, post = [], []
    for peer, s_row, r_row in ((down, 1, 0), (up, -2, -1)):
        if peer is None:
            continue
        snd = torch.stack([p[s_row, :] for p in planes]).contiguous()
        rcv = torch.empty_like(snd)
        pairs.append((peer, snd, rcv))
        post.append((rcv, r_row))
    _exchange(pairs)
    for rcv, r_row in post:
        for k, p in enumerate(planes):
            p[r_row, :] = rcv[k]
