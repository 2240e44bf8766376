"""Model / image loaders mirroring /root/reference/src/io.rs:24-90.

`load_wnn` reads the BTHOWeN-0g HDF5 model files with a minimal pure-Python HDF5 reader (the
reference links libhdf5, absent here): superblock v0, old-style groups, v1 object headers,
contiguous unfiltered datasets.  `synthetic_wnn` builds the same-shape
stand-in for the model file that is missing from the reference checkout (.MISSING_LARGE_BLOBS)."""
from __future__ import annotations

import struct

import numpy as np

from .wnn import Wnn


# ---- minimal HDF5 ---------------------------------------------------------------------------
class _H5:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n" and self.b[8] == 0, "unsupported HDF5 superblock"
        assert self.b[13] == 8 and self.b[14] == 8
        root = 56   # root symbol-table entry
        self.root_header = self.u64(root + 8)
        self.btree, self.heap = self.u64(root + 24), self.u64(root + 32)

    def u16(self, o):
        return struct.unpack_from("<H", self.b, o)[0]

    def u32(self, o):
        return struct.unpack_from("<I", self.b, o)[0]

    def u64(self, o):
        return struct.unpack_from("<Q", self.b, o)[0]

    def messages(self, addr):
        """v1 object header -> [(type, body_offset, size)] following continuation blocks."""
        assert self.b[addr] == 1
        nmsg = self.u16(addr + 2)
        size = self.u32(addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            o, sz = blocks.pop(0)
            end = o + sz
            while o + 8 <= end and len(out) < nmsg:
                t, s = self.u16(o), self.u16(o + 2)
                body = o + 8
                out.append((t, body, s))
                if t == 0x10:
                    blocks.append((self.u64(body), self.u64(body + 8)))
                o = body + s
        return out

    def children(self):
        """name -> object header address for the root group."""
        heap_data = self.u64(self.heap + 24)
        out = {}

        def walk(node):
            assert self.b[node:node + 4] == b"TREE"
            level, used = self.b[node + 5], self.u16(node + 6)
            o = node + 24
            for i in range(used):
                child = self.u64(o + 8 + 16 * i)
                if level > 0:
                    walk(child)
                else:
                    assert self.b[child:child + 4] == b"SNOD"
                    for s in range(self.u16(child + 6)):
                        e = child + 8 + 40 * s
                        no = heap_data + self.u64(e)
                        name = self.b[no:self.b.index(b"\0", no)].decode()
                        out[name] = self.u64(e + 8)
        walk(self.btree)
        return out

    def _dtype(self, o):
        cls = self.b[o] & 0x0F
        size = self.u32(o + 4)
        if cls == 0:
            signed = bool(self.b[o + 1] & 0x08)
            return np.dtype("<%s%d" % ("i" if signed else "u", size))
        if cls == 1:
            return np.dtype("<f%d" % size)
        if cls == 8:
            return self._dtype(o + 8)      # enum: base type
        raise ValueError("unsupported HDF5 datatype class %d" % cls)

    def attrs(self, addr):
        out = {}
        for t, body, _ in self.messages(addr):
            if t != 0x0C:
                continue
            assert self.b[body] == 1
            nsz, dsz, ssz = self.u16(body + 2), self.u16(body + 4), self.u16(body + 6)
            pad = lambda x: (x + 7) & ~7
            o = body + 8
            name = self.b[o:o + nsz].split(b"\0")[0].decode()
            o += pad(nsz)
            dt = self._dtype(o)
            o += pad(dsz) + pad(ssz)
            out[name] = np.frombuffer(self.b, dtype=dt, count=1, offset=o)[0]
        return out

    def dataset(self, addr):
        shape = dt = data = None
        for t, body, _ in self.messages(addr):
            if t == 0x01:
                rank = self.b[body + 1]
                shape = tuple(self.u64(body + 8 + 8 * i) for i in range(rank))
            elif t == 0x03:
                dt = self._dtype(body)
            elif t == 0x08:
                assert self.b[body] == 3 and self.b[body + 1] == 1, "only contiguous layout supported"
                data = self.u64(body + 2)
        count = int(np.prod(shape))
        return np.frombuffer(self.b, dtype=dt, count=count, offset=data).reshape(shape)


def load_wnn(path: str) -> Wnn:
    """src/io.rs:36-90."""
    h = _H5(path)
    a = h.attrs(h.root_header)
    ch = h.children()
    num_classes, num_inputs = int(a["num_classes"]), int(a["num_inputs"])
    bits_per_input, num_filter_inputs = int(a["bits_per_input"]), int(a["num_filter_inputs"])
    num_filter_entries, num_filter_hashes = int(a["num_filter_entries"]), int(a["num_filter_hashes"])
    p = int(a["p"])
    bloom = h.dataset(ch["bloom_filters"]).astype(bool)
    assert bloom.shape == (num_classes, num_inputs * bits_per_input // num_filter_inputs, num_filter_entries)
    width = int(np.float32(num_inputs) ** np.float32(0.5))
    thr = h.dataset(ch["binarization_thresholds"]).astype(np.float32)
    assert thr.shape == (width, width, bits_per_input)
    thr = np.minimum(np.maximum(np.ceil(thr * np.float32(255.0)), np.float32(0.0)), np.float32(256.0)).astype(np.uint16)
    order = h.dataset(ch["input_order"]).astype(np.uint64)
    assert order.shape == (num_inputs * bits_per_input,)
    return Wnn(num_classes, num_filter_entries, num_filter_hashes, num_filter_inputs, p, bloom, order, thr)


def load_grayscale_image(path: str) -> np.ndarray:
    """src/io.rs:24-33: first channel of the RGB-converted image, shape (height, width)."""
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))[:, :, 0].copy()


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    for q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def synthetic_wnn(num_classes=10, num_inputs=784, bits_per_input=6, num_filter_inputs=49, num_filter_entries=8192,
                  num_filter_hashes=4, density=0.12, zero_threshold_frac=0.3, seed=49_8192_4_6) -> Wnn:
    """Same-shape stand-in for models/model_49input_8192entry_4hash_6bpi.hdf5 (absent from the
    reference checkout; shape from models/readme.md:29-33 and src/lib.rs:51).  SURVEY.md 8(d)."""
    rng = np.random.default_rng(seed)
    l = num_filter_hashes * int(np.log2(num_filter_entries))
    p = (1 << (l + 1)) - 1
    while not _is_prime(p):
        p -= 2
    n_filters = num_inputs * bits_per_input // num_filter_inputs
    bloom = rng.random((num_classes, n_filters, num_filter_entries)) < density
    order = rng.permutation(num_inputs * bits_per_input).astype(np.uint64)
    width = int(round(num_inputs ** 0.5))
    t = np.sort(rng.random((width, width, bits_per_input)), axis=2)
    t[rng.random((width, width)) < zero_threshold_frac, 0] = 0.0
    thr = np.minimum(np.ceil(t.astype(np.float32) * np.float32(255.0)), 256).astype(np.uint16)
    return Wnn(num_classes, num_filter_entries, num_filter_hashes, num_filter_inputs, p, bloom, order, thr)


def synthetic_image(index: int, shape=(28, 28)) -> np.ndarray:
    """image i from default_rng(1000+i), 15 % non-zero pixels uniform in [1,255]."""
    rng = np.random.default_rng(1000 + index)
    img = np.zeros(shape, dtype=np.uint8)
    mask = rng.random(shape) < 0.15
    img[mask] = rng.integers(1, 256, size=int(mask.sum()), dtype=np.uint8)
    return img


# ---- wire formats shared with the Rust CLI (src/io.rs:137-207) -------------------------------------------------
# proof.json and circuit_params.json are serde_json renderings of structs defined IN the reference, so their
# layout is certain.  The SRS file is written by halo2_proofs' ParamsKZG::write (un-vendored): the layout below is
# the one of tag v2023_04_20 as recalled (SerdeFormat::RawBytes) and must be re-checked against a file written by
# the Rust CLI as soon as one is available.
def write_circuit_params(params: dict, path: str) -> None:
    """src/io.rs:150-152: serde_json of WnnCircuitParams (src/gadgets/wnn.rs:245-253), field order preserved."""
    import json
    keys = ["p", "l", "n_hashes", "bits_per_hash", "bits_per_filter", "n_classes"]
    with open(path, "w") as f:
        json.dump({k: int(params[k]) for k in keys}, f, separators=(",", ":"))


def read_circuit_params(path: str) -> dict:
    """src/io.rs:155-157"""
    import json
    with open(path) as f:
        d = json.load(f)
    return {k: int(d[k]) for k in ["p", "l", "n_hashes", "bits_per_hash", "bits_per_filter", "n_classes"]}


def write_proof_with_output(proof: bytes, output, path: str) -> None:
    """ProofWithOutput::write (src/io.rs:178-207): {"proof":[u8...],"output":[Fr...]}.  halo2curves 0.3.3 derives
    Serialize on `struct Fr([u64; 4])` (feature derive_serde, Cargo.toml:26-28), i.e. an Fr is the JSON array of its
    four little-endian Montgomery limbs -- the same limbs the C ABI uses."""
    import json
    from .bn254_host import to_limbs
    limbs = to_limbs([int(v) for v in output])
    with open(path, "w") as f:
        json.dump({"proof": list(proof), "output": [[int(x) for x in row] for row in limbs]}, f, separators=(",", ":"))


def read_proof_with_output(path: str):
    """ProofWithOutput::read: returns (proof bytes, [canonical ints])."""
    import json
    from .bn254_host import R_MOD, from_limbs
    with open(path) as f:
        d = json.load(f)
    out = np.array(d["output"], dtype=np.uint64).reshape(-1, 4)
    return bytes(d["proof"]), [int(v) for v in from_limbs(out, R_MOD)]


def write_srs(k: int, g: np.ndarray, g_lagrange: np.ndarray, g2_raw: bytes, s_g2_raw: bytes, path: str) -> None:
    """ParamsKZG::<Bn256>::write as used by src/io.rs:138-140 [UPSTREAM-RECALLED layout, RawBytes]:
    k (u32 LE) | g[0..n) | g_lagrange[0..n) | g2 | s_g2, G1Affine = x | y as 4 x u64 LE Montgomery limbs each (64 B),
    G2Affine = 128 B.  g / g_lagrange are the (n, 8) uint64 arrays the C ABI takes, so they are written verbatim."""
    n = 1 << k
    g = np.ascontiguousarray(g, dtype="<u8")
    gl = np.ascontiguousarray(g_lagrange, dtype="<u8")
    assert g.shape == (n, 8) and gl.shape == (n, 8) and len(g2_raw) == 128 and len(s_g2_raw) == 128
    with open(path, "wb") as f:
        f.write(int(k).to_bytes(4, "little"))
        f.write(g.tobytes())
        f.write(gl.tobytes())
        f.write(g2_raw)
        f.write(s_g2_raw)


def read_srs(path: str):
    """ParamsKZG::<Bn256>::read (src/io.rs:143-145), same layout: returns (k, g, g_lagrange, g2_raw, s_g2_raw).
    The prover needs g and g_lagrange only; the G2 points are carried through for the verifier."""
    with open(path, "rb") as f:
        k = int.from_bytes(f.read(4), "little")
        if not 1 <= k <= 28:
            raise ValueError("SRS file: k = %d out of range (not a RawBytes ParamsKZG file?)" % k)
        n = 1 << k
        g = np.frombuffer(f.read(64 * n), dtype="<u8").reshape(n, 8).astype(np.uint64)
        gl = np.frombuffer(f.read(64 * n), dtype="<u8").reshape(n, 8).astype(np.uint64)
        g2, s_g2 = f.read(128), f.read(128)
        if len(s_g2) != 128 or f.read(1):
            raise ValueError("SRS file: unexpected length for k = %d" % k)
    return k, g, gl, g2, s_g2


def encode_calldata(instances, proof: bytes) -> bytes:
    """snark_verifier::loader::evm::encode_calldata as called at src/eth.rs:114 / :211 [UPSTREAM-RECALLED]: every public
    input as a 32-byte big-endian canonical scalar, instance columns in order, followed by the proof bytes -- the
    calldata the generated EVM verifier takes."""
    from .bn254_host import R_MOD
    out = bytearray()
    for col in instances:
        for v in col:
            out += (int(v) % R_MOD).to_bytes(32, "big")
    return bytes(out) + bytes(proof)
