"""`Wnn`: host-side mirror of /root/reference/src/wnn.rs (same method names and argument meaning).

predict / get_circuit_params / get_circuit / mock_proof are host logic; generate_proving_key / proof drive the B200
backend through the C ABI (zg_b200.lib) and have no CPU path; verify_proof is host code in the same library
(verification is host work in the reference too)."""
from __future__ import annotations

import numpy as np


class Wnn:
    def __init__(self, num_classes, num_filter_entries, num_filter_hashes, num_filter_inputs, p,
                 bloom_filters, input_order, binarization_thresholds):
        self.num_classes = int(num_classes)
        self.num_filter_entries = int(num_filter_entries)
        self.num_filter_hashes = int(num_filter_hashes)
        self.num_filter_inputs = int(num_filter_inputs)
        self.p = int(p)
        self.bloom_filters = np.asarray(bloom_filters, dtype=bool)                  # (C, N, E)
        self.input_permutation = np.asarray(input_order, dtype=np.uint64)           # (inputs * bpi,)
        self.binarization_thresholds = np.asarray(binarization_thresholds, dtype=np.uint16)  # (w, h, bpi)

    # ---- src/wnn.rs:81-173 ---------------------------------------------------------------------
    def thermometer_encoding(self, image) -> np.ndarray:
        img = np.asarray(image).astype(np.uint16)
        t = self.binarization_thresholds
        # bit order: b, then i, then j
        return (img[None, :, :] >= np.transpose(t, (2, 0, 1))).reshape(-1)

    def mish_mash_hash(self, x: int) -> int:
        return (x * x * x % self.p) % (self.num_filter_entries ** self.num_filter_hashes)

    def encode_image(self, image):
        bits = self.thermometer_encoding(image)
        assert bits.shape[0] == self.input_permutation.shape[0]
        permuted = bits[self.input_permutation.astype(np.int64)]
        out = []
        nb = self.num_filter_inputs
        for k in range(0, len(permuted) - nb + 1, nb):
            v = 0
            for b in permuted[k:k + nb][::-1]:
                v = (v << 1) + int(b)
            out.append(v)
        return out

    def bloom_filter_lookup(self, bloom_array, filter_index: int) -> bool:
        h = self.mish_mash_hash(filter_index)
        e = self.num_filter_entries
        return all(bool(bloom_array[(h // e ** i) % e]) for i in range(self.num_filter_hashes))

    def predict(self, image):
        idx = self.encode_image(image)
        assert len(idx) == self.bloom_filters.shape[1]
        return [sum(int(self.bloom_filter_lookup(self.bloom_filters[c, f], x)) for f, x in enumerate(idx))
                for c in range(self.num_classes)]

    # ---- src/wnn.rs:175-195 --------------------------------------------------------------------
    def get_circuit_params(self) -> dict:
        bph = int(np.float32(np.log2(np.float32(self.num_filter_entries))))
        return {"p": self.p, "l": self.num_filter_hashes * bph, "n_hashes": self.num_filter_hashes,
                "bits_per_hash": bph, "bits_per_filter": self.num_filter_inputs,
                "n_classes": self.bloom_filters.shape[0]}

    def img_shape(self):
        return self.binarization_thresholds.shape[0], self.binarization_thresholds.shape[1]

    def get_circuit(self):
        from .plonk.gadgets import WnnCircuit
        return WnnCircuit(self.get_circuit_params(), self.bloom_filters, self.binarization_thresholds,
                          self.input_permutation)

    def synthesize(self, image, k: int):
        """Runs WnnCircuit::synthesize under the SimpleFloorPlanner: returns (circuit, assembly)."""
        from .plonk.circuit import Assembly, SimpleFloorPlanner
        circuit = self.get_circuit()
        asm = Assembly(circuit.cs, k)
        circuit.synthesize(SimpleFloorPlanner(asm), np.asarray(image))
        return circuit, asm

    def mock_proof(self, image, k: int):
        """src/wnn.rs:203-210: MockProver::run(...).assert_satisfied()."""
        from .plonk.mock import assert_satisfied
        outputs = self.predict(image)
        circuit, asm = self.synthesize(image, k)
        assert_satisfied(circuit.cs, asm, [outputs])

    # ---- src/wnn.rs:222-281: proving through the B200 backend ------------------------------------
    def generate_proving_key(self, ctx, params, k: int = None):
        """keygen_vk + keygen_pk on a dummy (all-zero) image: keys do not depend on the input."""
        from .prover import keygen
        k = params.k if k is None else k
        circuit, asm = self.synthesize(np.zeros(self.img_shape(), dtype=np.uint8), k)
        return keygen(ctx, params, circuit.cs, asm)

    def native_synthesizer(self):
        """zg_wnn_* (csrc/wnn_synth.cu): the same witness as `synthesize`, in C++ (about 3 ms instead of 0.5 s at k = 15)."""
        from .lib import NativeSynthesizer
        if getattr(self, "_native", None) is None:
            self._native = NativeSynthesizer(self)
        return self._native

    def proof(self, pk, params, image, rng=None, native: bool = True):
        """create_proof::<KZG<Bn256>, ProverGWC, _, _, EvmTranscript, _>; returns (proof bytes, outputs).
        `native` selects the C++ witness synthesis (default) or the Python front-end; both give the same columns.
        `rng` = None draws from an OS-seeded ChaCha20 stream like the reference's OsRng (src/wnn.rs:256)."""
        from .prover import create_proof, create_proof_limbs
        if native:
            from .bn254_host import to_limbs
            usable = (1 << pk.k) - (pk.cs.blinding_factors() + 1)
            cols, outputs = self.native_synthesizer().synthesize(image, pk.k, usable)
            params.load(pk.ctx)
            return create_proof_limbs(pk, cols, [to_limbs(outputs)], rng), outputs
        outputs = self.predict(image)
        _, asm = self.synthesize(image, pk.k)
        return create_proof(params, pk, asm.advice, [outputs], rng), outputs

    def verify_proof(self, proof: bytes, pk, params, outputs) -> bool:
        """src/wnn.rs:265-280: verify_proof::<KZG<Bn256>, VerifierGWC, _, EvmTranscript, SingleStrategy> against the
        claimed class scores.  Host code (csrc/verifier.cu: transcript, expression evaluation, one small MSM, two
        pairings); `pk` may be a ProvingKey or a VerifyingKey."""
        vk = pk.get_vk() if hasattr(pk, "get_vk") else pk
        return vk.verify(params, [list(outputs)], proof)
