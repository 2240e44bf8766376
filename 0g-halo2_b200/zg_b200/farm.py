"""Multi-GPU sharding of the proving path: one process per GPU, launched with
torchrun; `torch.distributed` is plumbing only.

  * prove_many      independent proofs (BASELINE configs[4]): image i -> rank i mod world; every rank holds the
                    full SRS + proving key; NO data-path collective -- rank 0 gathers the ~3.8 KB proofs.
  * join_communicator + Context.msm_sharded / set_distribution: the NCCL paths INSIDE the library (csrc/dist.cu):
                    one large MSM split by point range (configs[3]: rank g owns bases [g n/G, (g+1) n/G) and the
                    matching scalar slice; ncclAllGather of the G 96-byte partial sums, G - 1 additions on the device),
                    and one proof's commitment rounds spread over the ranks by column.  torch.distributed only
                    carries the 128-byte NCCL id.
  * sharded_msm     the same point-range split driven from the host with `torch.distributed.all_gather` (works on
                    gloo: the CPU tests of the N > 1 host logic use it).  EC addition is not an NCCL reduce op, so
                    this is an all_gather, never an all_reduce.
NTT / quotient / lookup stages do not shard at these sizes (an all-to-all of 32-byte elements costs more than
the single-GPU pass): replicas only.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np

from .bn254_host import Q_MOD, from_limbs


def join_communicator(ctx, dist=None) -> None:
    """Gives `ctx` the library-level NCCL communicator of this job: rank 0 creates the id, torch.distributed
    broadcasts its 128 bytes, every rank calls zg_comm_init.  No-op for a single process."""
    from . import lib as zl
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [zl.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(world, rank, box[0])


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """round-robin assignment: item i -> rank i mod world"""
    return list(range(rank, n_items, world))


def point_range(n: int, rank: int, world: int):
    """contiguous base range of rank `rank` for an n-point MSM (last rank takes the remainder)"""
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return lo, hi


def prove_many(images: Sequence, prove_one: Callable[[int, object], bytes] = None, dist=None,
               prove_batch: Callable[[List[int], List[object]], List[bytes]] = None) -> List[bytes] | None:
    """Every rank proves its share; rank 0 returns the proofs in image order (others return None).
    `prove_one(index, image)` is the per-GPU prover (Wnn.proof bound to this rank's context).  For throughput give
    every rank a `service.ProofService` (several proofs in flight per GPU) and pass `prove_batch` instead: it receives
    this rank's (index, image) pairs at once and returns their proofs in that order."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    idx = shard_indices(len(images), rank, world)
    if prove_batch is not None:
        mine = list(zip(idx, prove_batch(idx, [images[i] for i in idx])))
    else:
        mine = [(i, prove_one(i, images[i])) for i in idx]
    if world == 1:
        return [p for _, p in mine]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * len(images)
    for part in gathered:
        for i, p in part:
            out[i] = p
    return out


def _jac_add_host(p, q):
    """Jacobian addition on canonical ints (host, <= 8 points per MSM): the G-1 local adds after the gather."""
    m = Q_MOD
    if p[2] == 0:
        return q
    if q[2] == 0:
        return p
    z1z1, z2z2 = p[2] * p[2] % m, q[2] * q[2] % m
    u1, u2 = p[0] * z2z2 % m, q[0] * z1z1 % m
    s1, s2 = p[1] * q[2] * z2z2 % m, q[1] * p[2] * z1z1 % m
    if u1 == u2:
        if s1 != s2:
            return (1, 1, 0)
        a, b = p[0] * p[0] % m, p[1] * p[1] % m
        c = b * b % m
        d = 2 * ((p[0] + b) ** 2 - a - c) % m
        e = 3 * a % m
        x3 = (e * e - 2 * d) % m
        return (x3, (e * (d - x3) - 8 * c) % m, 2 * p[1] * p[2] % m)
    h, r = (u2 - u1) % m, (s2 - s1) % m
    hh = h * h % m
    hhh, v = h * hh % m, u1 * hh % m
    x3 = (r * r - hhh - 2 * v) % m
    return (x3, (r * (v - x3) - s1 * hhh) % m, p[2] * q[2] * h % m)


def combine_partials(partials: np.ndarray):
    """(G,12) uint64 Jacobian limbs -> affine (x, y) ints or None"""
    acc = (1, 1, 0)
    for row in np.asarray(partials, dtype=np.uint64).reshape(-1, 12):
        x, y, z = from_limbs(row.reshape(3, 4), Q_MOD)
        acc = _jac_add_host(acc, (x, y, z))
    if acc[2] == 0:
        return None
    zi = pow(acc[2], -1, Q_MOD)
    return (acc[0] * zi * zi % Q_MOD, acc[1] * zi * zi * zi % Q_MOD)


def sharded_msm(local_msm: Callable[[np.ndarray], np.ndarray], scalars: np.ndarray, dist=None, device=None):
    """`local_msm(scalar_slice) -> (12,) Jacobian limbs` runs on this rank's GPU over its resident base range.
    All ranks hold the full scalar vector (host) and take their slice; returns the affine result on every rank."""
    import torch
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    lo, hi = point_range(scalars.shape[0], rank, world)
    part = np.ascontiguousarray(local_msm(scalars[lo:hi]), dtype=np.uint64).reshape(12)
    if world == 1:
        return combine_partials(part[None])
    t = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    allp = np.stack([o.cpu().numpy().view(np.uint64) for o in outs])
    return combine_partials(allp)
