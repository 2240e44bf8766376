"""Throughput front-end: several proofs in flight on one GPU.

`Wnn::proof` (/root/reference/src/wnn.rs:232-262) proves one image at a time; a proof of this size leaves a B200 idle
during its latency-bound stretches (MSM tails, sorts, host round trips for the Fiat-Shamir challenges).  Contexts of the
backend are independent, so `ProofService` keeps `lanes` of them per GPU -- each with its own streams, per-proof
workspace and two sets of pinned witness buffers; the SRS window tables and the key's resident columns exist once
(`zg_srs_share`, `zg_pk_clone`) -- and drives every lane from its own host thread (the ctypes
calls release the GIL).  Image in, proof bytes out: witness synthesis is the native `zg_wnn_synthesize`, run one image
ahead of the proof on a helper thread per lane.

This is what `bench.py --inflight K --synth native` measures (profiles/README.md: 115 -> 183 proofs/s from 1 to 4 lanes
on the 1024-entry MNIST model) and what `farm.prove_many` runs on every rank of a multi-GPU job.  No CPU fallback.
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import lib as zl
from .bn254_host import to_limbs
from .prover import ParamsKZG, create_proof_limbs


class _Lane:
    def __init__(self, service: "ProofService", index: int, ctx: zl.Context):
        self.ctx = ctx
        self.params = ParamsKZG(service.params.k, service.params.g, service.params.g_lagrange)
        if index == 0 or not service.share:
            self.pk = service.wnn.generate_proving_key(ctx, self.params)
        else:
            # read-only state exists once per GPU: the SRS window tables and the key's resident columns of lane 0
            first = service.lanes[0]
            self.params.share(ctx, first.ctx)
            self.pk = first.pk.clone(ctx)
        n = 1 << self.pk.k
        # two sets of pinned advice buffers: the witness of the next image is synthesized (host, C++) into one while
        # zg_create_proof reads the other
        self.bufs = []
        self._pinned = []
        for _ in range(2):
            try:
                import torch
                pinned = [torch.empty((n, 4), dtype=torch.int64).pin_memory() for _ in range(6)]
                self._pinned.append(pinned)
                self.bufs.append([t.numpy().view(np.uint64) for t in pinned])
            except Exception:                  # torch is only used for pinned host memory here
                self.bufs.append([np.empty((n, 4), dtype=np.uint64) for _ in range(6)])
        self.cols = self.bufs[0]
        self.flip = 0
        self.usable = n - (self.pk.cs.blinding_factors() + 1)
        self.synth_pool = ThreadPoolExecutor(1)

    def close(self):
        self.synth_pool.shutdown(wait=True)
        self.pk.close()


class ProofService:
    """`lanes` independent provers for one model on one GPU.

    rng_factory(job_index) -> an RNG `create_proof_limbs` accepts; the default is a fresh ChaCha20 stream keyed from the
    operating system per proof (lib.ChaCha20Rng.from_os), the counterpart of the reference's OsRng.  Seeded XorShift
    streams are for reproducible tests only."""

    def __init__(self, wnn, params: ParamsKZG, device: int = 0, lanes: int = 4,
                 rng_factory: Optional[Callable[[int], object]] = None, streams: Optional[Sequence[int]] = None,
                 contexts: Optional[Sequence[zl.Context]] = None, share: bool = True):
        assert lanes >= 1
        self.wnn, self.params, self.device, self.share = wnn, params, device, share
        self.synth = wnn.native_synthesizer()
        self.rng_factory = rng_factory or self._os_rng
        self._owned_ctx = []
        self.lanes: List[_Lane] = []
        for i in range(lanes):
            if contexts is not None and i < len(contexts):
                ctx = contexts[i]                                  # the caller's context (kept open by close())
            else:
                ctx = zl.Context(device, None if streams is None else streams[i])
                self._owned_ctx.append(ctx)
            self.lanes.append(_Lane(self, i, ctx))

    @staticmethod
    def _os_rng(_job: int) -> zl.ChaCha20Rng:
        return zl.ChaCha20Rng.from_os()

    @property
    def vk(self):
        """fixed / permutation commitments and transcript_repr (identical on every lane)"""
        return self.lanes[0].pk

    def _prove_on(self, lane: _Lane, job: int, image) -> Tuple[bytes, List[int]]:
        cols, scores = self.synth.synthesize(image, lane.pk.k, lane.usable, out=lane.cols)
        return create_proof_limbs(lane.pk, cols, [to_limbs(scores)], self.rng_factory(job)), scores

    def prove(self, image) -> Tuple[bytes, List[int]]:
        """one proof on lane 0 (the latency path)"""
        return self._prove_on(self.lanes[0], 0, image)

    def prove_many(self, images: Sequence) -> List[Tuple[bytes, List[int]]]:
        """All images, `lanes` at a time; results in input order.  Every lane runs a two-stage pipeline: while
        zg_create_proof works on image i, the lane's helper thread synthesizes the witness of its next image into the
        other pinned buffer set (the ctypes calls release the GIL).  The first error of any lane is re-raised."""
        jobs: "queue.Queue[int]" = queue.Queue()
        for i in range(len(images)):
            jobs.put(i)
        out: List[Optional[Tuple[bytes, List[int]]]] = [None] * len(images)
        errors: List[BaseException] = []

        def work(lane: _Lane):
            def claim():
                try:
                    i = jobs.get_nowait()
                except queue.Empty:
                    return None
                buf = lane.bufs[lane.flip]
                lane.flip ^= 1
                return i, lane.synth_pool.submit(self.synth.synthesize, images[i], lane.pk.k, lane.usable, buf)
            cur = claim()
            while cur is not None and not errors:
                i, fut = cur
                nxt = claim()                    # the next witness is built while this proof runs
                try:
                    cols, scores = fut.result()
                    out[i] = (create_proof_limbs(lane.pk, cols, [to_limbs(scores)], self.rng_factory(i)), scores)
                except BaseException as e:      # surfaced to the caller below
                    errors.append(e)
                    if nxt is not None:
                        nxt[1].cancel()
                    return
                cur = nxt
        threads = [threading.Thread(target=work, args=(l,), daemon=True) for l in self.lanes[:max(1, min(len(self.lanes), len(images)))]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out  # type: ignore[return-value]

    def close(self):
        for l in self.lanes:
            l.close()
        for c in self._owned_ctx:
            c.close()
        self.lanes, self._owned_ctx = [], []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
