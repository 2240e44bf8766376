"""Multi-GPU parity checks of the NCCL paths inside the library (csrc/dist.cu).  Run with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist/run_dist_checks.py

  1. zg_msm_sharded: ONE 2^16-point MSM split by point range == the oracle's best_multiexp (host-pointer and device-pointer entry)
  2. column-distributed zg_create_proof (SPMD): every rank returns the SAME bytes as a single-GPU proof of the same inputs,
     and they equal the oracle's proof (tiny model, seeded RNG)
Exit status 0 on every rank = all checks passed (tests/test_gpu_dist.py launches this when >= 2 GPUs are visible)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "0g-halo2_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                      # plumbing only: carries the 128-byte NCCL id and the verdicts
    import bn254
    import cpu_ref
    import halo2_ref as H
    import zg_b200
    from zg_b200 import farm
    from zg_b200.io import load_grayscale_image, load_wnn
    from zg_b200.prover import ParamsKZG, create_proof, keygen
    ctx = zg_b200.Context(local)
    farm.join_communicator(ctx, dist)

    # ---- 1. point-range sharded MSM ---------------------------------------------------------------------------------
    logn = 16
    n = 1 << logn
    gen = bn254.g1_affine_to_limbs([bn254.G1_GEN])[0]
    bases = cpu_ref.g1_sequence(gen, n)
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 60) - 1)
    scalars = cpu_ref.fr_to_mont(raw)
    lo, hi = farm.point_range(n, rank, world)
    sub = zg_b200.Context(local)
    farm_ctx = ctx
    farm_ctx.srs_load((hi - lo).bit_length() - 1, np.ascontiguousarray(bases[lo:hi]), None)
    got = farm_ctx.msm_sharded(0, np.ascontiguousarray(scalars[lo:hi]))
    exp = cpu_ref.g1_to_affine(cpu_ref.best_multiexp(scalars, bases))
    assert (cpu_ref.g1_to_affine(got.reshape(1, 12)) == exp).all(), "rank %d: sharded MSM (host entry) differs from the oracle" % rank
    sc_dev = torch.from_numpy(np.ascontiguousarray(scalars[lo:hi]).view(np.int64)).cuda()
    out_dev = torch.zeros(12, dtype=torch.int64, device="cuda")
    farm_ctx.msm_sharded_dev(0, sc_dev.data_ptr(), hi - lo, hi - lo, 1, out_dev.data_ptr())
    farm_ctx.sync()
    got2 = out_dev.cpu().numpy().view(np.uint64)
    assert (cpu_ref.g1_to_affine(got2.reshape(1, 12)) == exp).all(), "rank %d: sharded MSM (device entry) differs" % rank

    # ---- 2. column-distributed proof -----------------------------------------------------------------------------------
    gold = os.path.join(ROOT, "tests", "golden")
    wnn = load_wnn(os.path.join(gold, "model_28input_256entry_1hash_1bpi.hdf5"))
    img = load_grayscale_image(os.path.join(gold, "example_image_7.png"))
    k = 14
    srs = H.Srs(k, 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8)
    params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
    outs = wnn.predict(img)
    _, asm = wnn.synthesize(img, k)
    seed = bytes(range(16))
    # single-GPU proof on a context without a communicator
    circ_s, asm_s = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    pk_s = keygen(sub, ParamsKZG(k, srs.g, srs.g_lagrange), circ_s.cs, asm_s)
    single = create_proof(ParamsKZG(k, srs.g, srs.g_lagrange), pk_s, asm.advice, [outs], zg_b200.lib.XorShift.from_seed(seed))
    # distributed
    circ_d, asm_d = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    pk_d = keygen(ctx, params, circ_d.cs, asm_d)
    ctx.set_distribution(1)
    for _ in range(2):                                     # twice: the cached lookup tables and workspaces are reused
        distributed = create_proof(params, pk_d, asm.advice, [outs], zg_b200.lib.XorShift.from_seed(seed))
        assert distributed == single, "rank %d: column-distributed proof differs from the single-GPU proof" % rank
    assert pk_d.get_vk().verify(params, [outs], distributed)
    if rank == 0:
        circ0, asm0 = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
        opk = H.keygen(srs, circ0.cs, asm0)
        oproof = H.create_proof(srs, opk, asm.advice, [outs], H.XorShiftRng(seed))
        assert oproof == distributed, "distributed proof differs from the oracle's"
    ctx.set_distribution(0)
    ctx.comm_destroy()
    dist.barrier()
    if rank == 0:
        print("dist checks ok on %d GPUs: sharded MSM == oracle, column-distributed proof == single-GPU proof == oracle" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
