"""N > 1 host logic on CPU: world_size-2 gloo process groups exercise the proof farm's sharding / gather and
the point-range MSM's all_gather + local EC adds (the per-GPU compute is stubbed by the CPU oracle here --
the sharding code itself never touches a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bn254
import cpu_ref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zg_b200 import farm
    # 1. proof farm: 5 "images", prover stub returns bytes naming (rank, index)
    images = list(range(5))
    proofs = farm.prove_many(images, lambda i, img: b"proof-%d-by-%d" % (i, rank), dist)
    if rank == 0:
        assert proofs == [b"proof-%d-by-%d" % (i, i % world) for i in range(5)]
    else:
        assert proofs is None
    # 1b. the batch hook (one ProofService call per rank) gives the same placement
    batched = farm.prove_many(images, dist=dist, prove_batch=lambda idx, imgs: [b"proof-%d-by-%d" % (i, rank) for i in idx])
    assert batched == proofs
    # 2. point-range MSM: each rank owns half of the bases; partials all-gathered; result identical on all ranks
    n = 64
    gen = bn254.g1_affine_to_limbs([bn254.G1_GEN])[0]
    bases = cpu_ref.g1_sequence(gen, n)
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 60) - 1)
    sc = cpu_ref.fr_to_mont(raw)
    lo, hi = farm.point_range(n, rank, world)
    got = farm.sharded_msm(lambda s: cpu_ref.best_multiexp(s, bases[lo:hi]), sc, dist)
    exp = bn254.g1_affine_from_limbs(cpu_ref.g1_to_affine(cpu_ref.best_multiexp(sc, bases)))[0]
    assert got == exp
    ret[rank] = 1
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()      # no fork() from a process that already runs OpenMP threads
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 1, 1: 1}


def test_sharding_helpers():
    from zg_b200 import farm
    assert farm.shard_indices(10, 1, 4) == [1, 5, 9]
    assert [farm.point_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    # combine_partials == group sum
    pts = [bn254.g1_mul(bn254.G1_GEN, k) for k in (3, 5, 11)]
    jac = np.zeros((3, 12), dtype=np.uint64)
    for i, p in enumerate(pts):
        jac[i, :8] = bn254.g1_affine_to_limbs([p])[0]
        jac[i, 8:] = bn254.ints_to_limbs([1], bn254.Q_MOD)[0]
    assert farm.combine_partials(jac) == bn254.g1_mul(bn254.G1_GEN, 19)
