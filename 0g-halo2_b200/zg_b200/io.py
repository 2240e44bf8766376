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


# ---- proving / verifying key files (src/io.rs:159-176) ------------------------------------------------------------
# `pk.write(writer, RawBytes)` / `pk.get_vk().write(writer, RawBytes)` of halo2_proofs v2023_04_20 (un-vendored;
# layout [UPSTREAM-RECALLED] from plonk.rs / permutation.rs / poly.rs / helpers.rs of that tag, to be re-checked against
# a file written by the Rust CLI):
#   VerifyingKey : k (u32 BE) | num_fixed (u32 BE) | fixed commitments | permutation commitments (no count: one per
#                  permutation column of the constraint system) | one bit per row per selector, packed LSB first,
#                  ceil(n/8) bytes per selector (the selector activations BEFORE compression)
#   ProvingKey   : VerifyingKey | l0 | l_last | l_active_row (Polynomial) | fixed_values | fixed_polys | fixed_cosets
#                  (slices) | permutation.{permutations, polys, cosets} (slices)
#   Polynomial   : len (u32 BE) | len field elements;   slice: count (u32 BE) | polynomials
#   RawBytes     : Fr = four u64 LE Montgomery limbs (32 B); G1Affine = x | y, Fq likewise (64 B) -- the limbs the
#                  C ABI uses, so columns are written verbatim.
# The extended forms in the file live on halo2's coset (zeta * <omega_ext>, 2^extended_k points), not on this
# backend's internal domain: `zg_b200.prover.export_proving_key` rebuilds them through the C ABI.
def _u32be(v: int) -> bytes:
    return int(v).to_bytes(4, "big")


def _read_exact(f, nbytes: int) -> bytes:
    b = f.read(nbytes)
    if len(b) != nbytes:
        raise ValueError("key file truncated")
    return b


def _write_poly(f, col) -> None:
    a = np.ascontiguousarray(col, dtype="<u8").reshape(-1, 4)
    f.write(_u32be(a.shape[0]))
    f.write(a.tobytes())


def _read_poly(f, expect_len=None) -> np.ndarray:
    n = int.from_bytes(_read_exact(f, 4), "big")
    if expect_len is not None and n != expect_len:
        raise ValueError("key file: polynomial of %d elements where %d were expected" % (n, expect_len))
    return np.frombuffer(_read_exact(f, 32 * n), dtype="<u8").reshape(n, 4).astype(np.uint64)


def _write_poly_slice(f, cols) -> None:
    f.write(_u32be(len(cols)))
    for c in cols:
        _write_poly(f, c)


def _read_poly_slice(f, expect_count=None, expect_len=None) -> list:
    cnt = int.from_bytes(_read_exact(f, 4), "big")
    if expect_count is not None and cnt != expect_count:
        raise ValueError("key file: %d polynomials where %d were expected" % (cnt, expect_count))
    return [_read_poly(f, expect_len) for _ in range(cnt)]


def write_vk(f, k: int, fixed_commitments, perm_commitments, selectors) -> None:
    """VerifyingKey::write(RawBytes).  Commitments as (count, 8) uint64 limb arrays, selectors as lists of bools."""
    fc = np.ascontiguousarray(fixed_commitments, dtype="<u8").reshape(-1, 8)
    pc = np.ascontiguousarray(perm_commitments, dtype="<u8").reshape(-1, 8)
    f.write(_u32be(k))
    f.write(_u32be(fc.shape[0]))
    f.write(fc.tobytes())
    f.write(pc.tobytes())
    for sel in selectors:
        bits = np.asarray(sel, dtype=bool)
        assert bits.shape[0] == 1 << k
        f.write(np.packbits(bits, bitorder="little").tobytes())


def read_vk(f, num_perm_columns: int, num_selectors: int) -> dict:
    """VerifyingKey::read(RawBytes): the caller supplies what the Rust side takes from the circuit (`params` ->
    configure): the number of permutation columns and of selectors.  Returns k, commitments, selector activations."""
    k = int.from_bytes(_read_exact(f, 4), "big")
    if not 1 <= k <= 28:
        raise ValueError("key file: k = %d out of range (not a RawBytes key file?)" % k)
    nf = int.from_bytes(_read_exact(f, 4), "big")
    if nf > 1 << 16:
        raise ValueError("key file: implausible number of fixed columns")
    fc = np.frombuffer(_read_exact(f, 64 * nf), dtype="<u8").reshape(nf, 8).astype(np.uint64)
    pc = np.frombuffer(_read_exact(f, 64 * num_perm_columns), dtype="<u8").reshape(num_perm_columns, 8).astype(np.uint64)
    n = 1 << k
    sels = []
    for _ in range(num_selectors):
        raw = np.frombuffer(_read_exact(f, (n + 7) // 8), dtype=np.uint8)
        sels.append(np.unpackbits(raw, bitorder="little")[:n].astype(bool).tolist())
    return {"k": k, "fixed_commitments": fc, "perm_commitments": pc, "selectors": sels}


def write_pk(f, vk: dict, l0, l_last, l_active_row, fixed_values, fixed_polys, fixed_cosets, perm_values, perm_polys,
             perm_cosets) -> None:
    """ProvingKey::write(RawBytes); `vk` as returned by read_vk / assembled by export_proving_key."""
    write_vk(f, vk["k"], vk["fixed_commitments"], vk["perm_commitments"], vk["selectors"])
    for p in (l0, l_last, l_active_row):
        _write_poly(f, p)
    for sl in (fixed_values, fixed_polys, fixed_cosets, perm_values, perm_polys, perm_cosets):
        _write_poly_slice(f, sl)


def read_pk(f, num_perm_columns: int, num_selectors: int) -> dict:
    """ProvingKey::read(RawBytes) (src/io.rs:166-170).  Lengths are cross-checked against k."""
    d = read_vk(f, num_perm_columns, num_selectors)
    n = 1 << d["k"]
    d["l0"], d["l_last"], d["l_active_row"] = _read_poly(f), _read_poly(f), _read_poly(f)
    ext_n = d["l0"].shape[0]
    if ext_n < n or ext_n & (ext_n - 1) or d["l_last"].shape[0] != ext_n or d["l_active_row"].shape[0] != ext_n:
        raise ValueError("key file: inconsistent extended-domain size")
    nf = d["fixed_commitments"].shape[0]
    d["fixed_values"] = _read_poly_slice(f, nf, n)
    d["fixed_polys"] = _read_poly_slice(f, nf, n)
    d["fixed_cosets"] = _read_poly_slice(f, nf, ext_n)
    d["perm_values"] = _read_poly_slice(f, num_perm_columns, n)
    d["perm_polys"] = _read_poly_slice(f, num_perm_columns, n)
    d["perm_cosets"] = _read_poly_slice(f, num_perm_columns, ext_n)
    if f.read(1):
        raise ValueError("key file: trailing bytes")
    return d
