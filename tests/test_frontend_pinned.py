"""The reference's own circuit-level tests, reproduced against the host front-end (zg_b200.plonk):

  * MockProver on the checked-in models at the documented k   /root/reference/tests/integration_test.rs:5-62
  * the 4x3 toy circuit with hand-derived hashes                /root/reference/src/gadgets/wnn.rs:401-494
  * the gadget KATs and bad-witness rejections                  src/gadgets/{hash,range_check,greater_than,bits2num}.rs,
                                                                src/gadgets/bloom_filter{,/array_lookup,/byte_selector,/bit_selector}.rs
and the cross-check of the product's constraint-system bookkeeping (degree, blinding factors, selector compression,
query tables, substituted expressions) against the oracle's independent restatement (oracle/circuit_ref.py).
CPU only."""
import os

import numpy as np
import pytest

import circuit_ref
from zg_b200.io import load_grayscale_image, load_wnn, synthetic_wnn
from zg_b200.plonk import gadgets as G
from zg_b200.plonk.circuit import Assembly, ConstraintSystem, SimpleFloorPlanner
from zg_b200.plonk.mock import finalize_fixed, verify
from zg_b200.wnn import Wnn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IMG = os.path.join(GOLD, "example_image_7.png")
MODELS = {  # src/lib.rs:48-51 (checked_in_test_data)
    "tiny": ("model_28input_256entry_1hash_1bpi.hdf5", 14),
    "small": ("model_28input_1024entry_2hash_2bpi.hdf5", 15),
    "medium": ("model_28input_2048entry_2hash_3bpi.hdf5", 15),
}


def run_mock(cs, k, synth, instances):
    """MockProver::run(k, circuit, instances).verify() -> list of failures"""
    asm = Assembly(cs, k)
    synth(SimpleFloorPlanner(asm))
    return verify(cs, asm, instances)


# ---- tests/integration_test.rs:5-62 -----------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny", "small", "medium"])
def test_mock_proof_mnist(name):
    fname, k = MODELS[name]
    wnn = load_wnn(os.path.join(GOLD, fname))
    wnn.mock_proof(load_grayscale_image(IMG), k)


def test_mock_proof_mnist_large_standin():
    """mock_proof_mnist_large (integration_test.rs:56-62) on the synthetic same-shape stand-in (the real file is absent
    upstream, .MISSING_LARGE_BLOBS) at the documented k = 17 (src/lib.rs:51)."""
    wnn = synthetic_wnn()
    wnn.mock_proof(load_grayscale_image(IMG), 17)


def test_mock_proof_rejects_wrong_output():
    fname, k = MODELS["tiny"]
    wnn = load_wnn(os.path.join(GOLD, fname))
    img = load_grayscale_image(IMG)
    out = wnn.predict(img)
    circ, asm = wnn.synthesize(img, k)
    out[3] += 1
    assert verify(circ.cs, asm, [out])


# ---- src/gadgets/wnn.rs:401-494 -------------------------------------------------------------------------------------
TOY_PARAMS = {"p": 2097143, "l": 20, "n_hashes": 2, "bits_per_hash": 10, "bits_per_filter": 12, "n_classes": 2}


def toy_wnn():
    image = np.array([[70, 100, 150], [20, 110, 200], [27, 50, 211], [200, 100, 3]], dtype=np.uint8)
    thresholds = np.array([[[50, 150], [0, 50], [200, 256]],
                           [[10, 80], [100, 200], [50, 150]],
                           [[0, 100], [100, 200], [0, 100]],
                           [[0, 100], [100, 200], [0, 100]]], dtype=np.uint16)
    perm = np.array(list(range(6, 24)) + list(range(6)), dtype=np.uint64)
    bloom = np.zeros((2, 2, 1024), dtype=bool)
    for c, f, e in [(0, 0, 966), (0, 0, 805), (0, 1, 494), (1, 0, 966), (1, 0, 805), (1, 1, 494), (1, 1, 46)]:
        bloom[c, f, e] = True
    return image, thresholds, perm, bloom


def test_toy_circuit_hand_derived_values():
    image, thresholds, perm, bloom = toy_wnn()
    # host model (src/wnn.rs:81-173) reproduces the comment block of the reference test
    wnn = Wnn(2, 1024, 2, 12, TOY_PARAMS["p"], bloom, perm, thresholds)
    bits = wnn.thermometer_encoding(image).astype(int).tolist()
    assert bits == [1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 1, 1, 0, 0]
    assert wnn.encode_image(image) == [2237, 3788]
    assert [wnn.mish_mash_hash(x) for x in (2237, 3788)] == [825286, 47598]
    assert (825286 % 1024, 825286 // 1024, 47598 % 1024, 47598 // 1024) == (966, 805, 494, 46)
    assert wnn.predict(image) == [1, 2]
    assert wnn.get_circuit_params() == TOY_PARAMS
    # the circuit: the same intermediate values appear as witness cells, and MockProver is satisfied at k = 13
    circ = G.WnnCircuit(TOY_PARAMS, bloom, thresholds, perm)
    seen = {"joint": [], "hash": []}
    orig_hash = circ.hash.hash

    def spy(lay, cell):
        out = orig_hash(lay, cell)
        seen["joint"].append(cell.value)
        seen["hash"].append(out.value)
        return out
    circ.hash.hash = spy
    asm = Assembly(circ.cs, 13)
    results = circ.synthesize(SimpleFloorPlanner(asm), image)
    assert seen == {"joint": [2237, 3788], "hash": [825286, 47598]}
    assert [c.value for c in results] == [1, 2]
    assert verify(circ.cs, asm, [[1, 2]]) == []
    assert verify(circ.cs, asm, [[1, 1]])            # a wrong public output breaks the instance copy constraint
    # a tampered witness cell breaks a gate: flip the second class score accumulator
    col, row = results[1].column.index, results[1].row
    asm.advice[col][row] = 3
    assert any("accumulate_responses" in f for f in verify(circ.cs, asm, [[1, 3]]))


# ---- src/gadgets/hash.rs:222-372 ------------------------------------------------------------------------------------
def hash_circuit(x):
    cs = ConstraintSystem()
    inp, quotient, remainder, msb, hsh = (cs.advice_column() for _ in range(5))
    cs.enable_constant(cs.fixed_column())
    instance = cs.instance_column()
    cs.enable_equality(instance)
    for c in (inp, quotient, remainder, msb, hsh):
        cs.enable_equality(c)
    table = cs.lookup_table_column()
    rc = G.RangeCheck(cs, inp, table)
    chip = G.Hash(cs, inp, quotient, remainder, msb, hsh, rc, p=11, l=3, n_bits=8)

    def synth(lay):
        cell = lay.assign_region("input", lambda r: r.assign_advice(inp, 0, x))
        G.load_bytes_column(lay, table)
        out = chip.hash(lay, cell)
        lay.constrain_instance(out, instance, 0)
    return cs, synth


@pytest.mark.parametrize("x,expected", [(2, 0), (4, 1), (42, 3), (255, 0)])   # (x^3 % 11) % 8
def test_hash_kat(x, expected):
    cs, synth = hash_circuit(x)
    assert run_mock(cs, 9, synth, [[expected]]) == []


def test_hash_rejects_wrong_output():
    cs, synth = hash_circuit(42)
    assert run_mock(cs, 9, synth, [[2]])


# ---- src/gadgets/range_check.rs:150-290 -------------------------------------------------------------------------------
def le_circuit(x, y):
    cs = ConstraintSystem()
    adv = cs.advice_column()
    table = cs.lookup_table_column()
    constants = cs.fixed_column()
    cs.enable_equality(adv)
    cs.enable_constant(constants)
    rc = G.RangeCheck(cs, adv, table)

    def synth(lay):
        cell = lay.assign_region("x", lambda r: r.assign_advice(adv, 0, x))
        G.load_bytes_column(lay, table)
        rc.le_constant(lay, cell, y)
    return cs, synth


@pytest.mark.parametrize("x,y", [(1023, 1023), (1022, 1023), (4, 9), (0, 0xFFABCDEF)])
def test_le_constant_satisfied(x, y):
    cs, synth = le_circuit(x, y)
    assert run_mock(cs, 9, synth, []) == []


def test_le_constant_rejects_greater():          # test_le_greater_10bit: 1024 <= 1023 must fail
    cs, synth = le_circuit(1024, 1023)
    assert run_mock(cs, 9, synth, [])


# ---- src/gadgets/greater_than.rs:200-330 --------------------------------------------------------------------------------
def gt_circuit(x, y):
    cs = ConstraintSystem()
    instance = cs.instance_column()
    cx, cy, diff, is_gt = (cs.advice_column() for _ in range(4))
    constants = cs.fixed_column()
    byte_column = cs.lookup_table_column()
    cs.enable_constant(constants)
    cs.enable_equality(instance)
    for c in (cx, cy, diff, is_gt):
        cs.enable_equality(c)
    rc = G.RangeCheck(cs, cx, byte_column)
    chip = G.GreaterThan(cs, cx, cy, diff, is_gt, rc)

    def synth(lay):
        G.load_bytes_column(lay, byte_column)
        _, gt = chip.witness(lay, x, y)
        lay.constrain_instance(gt, instance, 0)
    return cs, synth


@pytest.mark.parametrize("x,y,out", [(129, 64, 1), (64, 129, 0), (64, 64, 0)])
def test_greater_than(x, y, out):
    cs, synth = gt_circuit(x, y)
    assert run_mock(cs, 9, synth, [[out]]) == []


def test_greater_than_rejects_x_too_large():      # test_x_too_large: x = 256 is not a byte
    cs, synth = gt_circuit(256, 64)
    assert run_mock(cs, 9, synth, [[0]])


# ---- src/gadgets/bits2num.rs:130-265 ---------------------------------------------------------------------------------------
def test_bits2num_le():
    cs = ConstraintSystem()
    inp, acc = cs.advice_column(), cs.advice_column()
    constants = cs.fixed_column()
    pub = cs.instance_column()
    cs.enable_equality(pub)
    cs.enable_equality(acc)
    cs.enable_equality(inp)
    cs.enable_constant(constants)
    chip = G.Bits2Num(cs, inp, acc)

    def synth(lay):
        bits = [lay.assign_region("input bit %d" % i, lambda r, i=i, b=b: r.assign_advice(inp, i, int(b)))
                for i, b in enumerate([True, False, True, False])]
        lay.constrain_instance(chip.convert_le(lay, bits), pub, 0)
    assert run_mock(cs, 5, synth, [[5]]) == []
    cs2 = ConstraintSystem()
    inp, acc = cs2.advice_column(), cs2.advice_column()
    constants = cs2.fixed_column()
    pub = cs2.instance_column()
    for c in (pub, acc, inp):
        cs2.enable_equality(c)
    cs2.enable_constant(constants)
    chip = G.Bits2Num(cs2, inp, acc)
    assert run_mock(cs2, 5, synth, [[10]])        # 0b0101 little-endian is 5, not 10


# ---- src/gadgets/bloom_filter/bit_selector.rs:165-300 ------------------------------------------------------------------------
@pytest.mark.parametrize("index,out", [(0, 1), (1, 1), (7, 0)])
def test_bit_selector(index, out):
    cs = ConstraintSystem()
    byte, idx, bit = (cs.advice_column() for _ in range(3))
    instance = cs.instance_column()
    cs.enable_equality(instance)
    for c in (byte, idx, bit):
        cs.enable_equality(c)
    chip = G.BitSelector(cs, byte, idx, bit)

    def synth(lay):
        def body(r):
            return r.assign_advice(byte, 0, 0b11111110), r.assign_advice(idx, 0, index)
        bc, ic = lay.assign_region("inputs", body)
        chip.load(lay)
        lay.constrain_instance(chip.select(lay, bc, ic), instance, 0)
    assert run_mock(cs, 12, synth, [[out]]) == []


# ---- src/gadgets/bloom_filter/byte_selector.rs:350-535 -----------------------------------------------------------------------
@pytest.mark.parametrize("word,index,num_bytes,out", [(0xAB, 0, 1, 0xAB), (0xABCDEF, 0, 3, 0xAB), (0xABCDEF, 1, 3, 0xCD)])
def test_byte_selector(word, index, num_bytes, out):
    cs = ConstraintSystem()
    cols = [cs.advice_column() for _ in range(6)]
    instance = cs.instance_column()
    constants = cs.fixed_column()
    table = cs.lookup_table_column()
    cs.enable_equality(instance)
    for c in cols:
        cs.enable_equality(c)
    cs.enable_constant(constants)
    chip = G.ByteSelector(cs, *cols, table)

    def synth(lay):
        def body(r):
            return r.assign_advice(cols[0], 0, word), r.assign_advice(cols[1], 0, index)
        wc, ic = lay.assign_region("inputs", body)
        G.load_bytes_column(lay, table)
        lay.constrain_instance(chip.select(lay, wc, ic, num_bytes), instance, 0)
    assert run_mock(cs, 9, synth, [[out]]) == []


# ---- src/gadgets/bloom_filter/array_lookup.rs:470-700 ---------------------------------------------------------------------------
WORDS = [0x1122334455667788, 0x99AABBCCDDEEFF00, 0xBABABABABABABABA, 0x0123456789ABCDEF,
         0x1111111111111111, 0x2222222222222222, 0x3333333333333333, 0x4444444444444444]


@pytest.mark.parametrize("x,bloom_index,out", [
    (0b_01_001_101_00_111_000, 0, [WORDS[1], 0b001, 0b101, WORDS[0], 0b111, 0b000]),
    (0b_11_001_101_11_111_000, 0, [WORDS[3], 0b001, 0b101, WORDS[3], 0b111, 0b000]),
    (0b_01_001_101_00_111_000, 1, [WORDS[5], 0b001, 0b101, WORDS[4], 0b111, 0b000]),
])
def test_array_lookup(x, bloom_index, out):
    bits = np.array([[(w >> (63 - i)) & 1 for w in WORDS[4 * a:4 * a + 4] for i in range(64)] for a in range(2)], dtype=bool)
    cs = ConstraintSystem()
    instance = cs.instance_column()
    adv = [cs.advice_column() for _ in range(5)]
    for a in adv:
        cs.enable_equality(a)
    cs.enable_equality(instance)
    cs.enable_constant(cs.fixed_column())
    chip = G.ArrayLookup(cs, *adv, n_hashes=2, bits_per_hash=8, word_index_bits=2)
    chip.set_arrays(bits)

    def synth(lay):
        cell = lay.assign_region("input", lambda r: r.assign_advice(adv[0], 0, x))
        chip.load(lay)
        res = chip.lookup(lay, cell, bloom_index)
        assert len(res) == 2
        for i, (word, byte_index, bit_index) in enumerate(res):
            lay.constrain_instance(word, instance, 3 * i)
            lay.constrain_instance(byte_index, instance, 3 * i + 1)
            lay.constrain_instance(bit_index, instance, 3 * i + 2)
    assert run_mock(cs, 10, synth, [out]) == []


# ---- src/gadgets/bloom_filter.rs:195-385 ------------------------------------------------------------------------------------------
def bloom_case(arrays, x, out):
    cs = ConstraintSystem()
    instance = cs.instance_column()
    adv = [cs.advice_column() for _ in range(6)]
    for a in adv:
        cs.enable_equality(a)
    cs.enable_equality(instance)
    cs.enable_constant(cs.fixed_column())
    chip = G.BloomFilter(cs, adv, n_hashes=2, bits_per_hash=10)
    chip.array.set_arrays(arrays)

    def synth(lay):
        cell = lay.assign_region("input", lambda r: r.assign_advice(adv[0], 0, x))
        chip.load(lay)
        lay.constrain_instance(chip.bloom_lookup(lay, cell, 0), instance, 0)
    return run_mock(cs, 14, synth, [[out]])


def test_bloom_filter_all_positive():
    assert bloom_case(np.ones((1, 1024), dtype=bool), 8, 1) == []


def test_bloom_filter_all_negative():
    assert bloom_case(np.zeros((1, 1024), dtype=bool), 8, 0) == []


def test_bloom_filter_index_1_2():
    a = np.zeros((1, 1024), dtype=bool)
    a[0, 1] = a[0, 2] = True
    assert bloom_case(a, 0b0000000001_0000000010, 1) == []
    b = np.zeros((1, 1024), dtype=bool)
    b[0, 0] = b[0, 2] = True
    assert bloom_case(b, 0b0000000001_0000000010, 0) == []
    assert bloom_case(b, 0b0000000001_0000000010, 1)          # claiming a positive response is rejected


# ---- product bookkeeping == the oracle's independent restatement (oracle/circuit_ref.py) ------------------------------------------
def _compare_with_refcs(circ, asm):
    cs = circ.cs
    ref = circuit_ref.RefCS.from_frontend(cs)                      # before the product compresses anything
    assert ref.degree() == cs.degree() and ref.blinding_factors() == cs.blinding_factors()
    nf0 = ref.num_fixed
    ref_new = ref.compress_selectors([list(a) for a in asm.selectors])
    fixed = finalize_fixed(cs, asm)                                  # the product's compression (mutates cs)
    assert ref.num_fixed == cs.num_fixed == len(fixed)
    assert fixed[nf0:] == ref_new
    assert ref.queries == {k: [tuple(q) for q in v] for k, v in cs.queries.items()}
    assert ref.degree() == cs.degree() and ref.blinding_factors() == cs.blinding_factors()
    for (_, rp), g in zip(ref.gate_polys, cs.gates):
        assert rp == [circuit_ref.frontend_expr_tuple(p) for p in g.polys]
    for (_, ri, rt), l in zip(ref.lookup_exprs, cs.lookups):
        assert ri == [circuit_ref.frontend_expr_tuple(e) for e in l.inputs]
        assert rt == [circuit_ref.frontend_expr_tuple(e) for e in l.tables]
    return ref


def test_refcs_matches_product_on_toy_circuit():
    image, thresholds, perm, bloom = toy_wnn()
    circ = G.WnnCircuit(TOY_PARAMS, bloom, thresholds, perm)
    asm = Assembly(circ.cs, 13)
    circ.synthesize(SimpleFloorPlanner(asm), image)
    ref = _compare_with_refcs(circ, asm)
    # degree 6 (so the extended domain is 8n), 6 advice / 1 instance columns, 4 lookups and the 12 gate
    # polynomials its table A.2 lists (4 + 1 + 1 + 1 + 1 + 2 + 1 + 1; the heading's "13" is a miscount of that table)
    assert ref.degree() == 6 and ref.num_advice == 6 and ref.num_instance == 1
    assert len(ref.lookup_exprs) == 4 and sum(len(p) for _, p in ref.gate_polys) == 12


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_refcs_matches_product_on_models(name):
    fname, k = MODELS[name]
    wnn = load_wnn(os.path.join(GOLD, fname))
    circ, asm = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    ref = _compare_with_refcs(circ, asm)
    assert ref.num_fixed == 16                      # SURVEY A.1 ("16 fixed columns after compression")


@pytest.mark.parametrize("which", ["hash", "greater_than", "bloom"])
def test_refcs_matches_product_on_gadget_circuits(which):
    """the same cross-check on three gadget test circuits: other selector mixes, other degrees, other query tables"""
    class Holder:
        pass
    if which == "hash":
        cs, synth = hash_circuit(42)
        k = 9
    elif which == "greater_than":
        cs, synth = gt_circuit(129, 64)
        k = 9
    else:
        cs = ConstraintSystem()
        instance = cs.instance_column()
        adv = [cs.advice_column() for _ in range(6)]
        for a in adv:
            cs.enable_equality(a)
        cs.enable_equality(instance)
        cs.enable_constant(cs.fixed_column())
        chip = G.BloomFilter(cs, adv, n_hashes=2, bits_per_hash=10)
        chip.array.set_arrays(np.ones((1, 1024), dtype=bool))

        def synth(lay):
            cell = lay.assign_region("input", lambda r: r.assign_advice(adv[0], 0, 8))
            chip.load(lay)
            lay.constrain_instance(chip.bloom_lookup(lay, cell, 0), instance, 0)
        k = 14
    asm = Assembly(cs, k)
    synth(SimpleFloorPlanner(asm))
    h = Holder()
    h.cs = cs
    _compare_with_refcs(h, asm)
