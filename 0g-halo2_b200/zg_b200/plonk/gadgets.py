"""Host-side mirror of zero_g's circuit: `WnnChip` and its sub-chips.

Each class restates the `configure` (constraint system) and the region-assignment code of one
gadget of the reference, in the same call order, so that the resulting columns, selector rows,
constants and copy-constraint order match what halo2's SimpleFloorPlanner would produce for the
Rust circuit.  Citations are to /root/reference/src/gadgets/**.  Values are Python ints mod r.

This is the workload definition of the proving path; witness
synthesis stays on the host (BASELINE.json north_star)."""
from __future__ import annotations

import math
from typing import List

from ..bn254_host import R_MOD
from .circuit import (Cell, Column, Const, ConstraintSystem, Expr, Region, Selector, SimpleFloorPlanner,
                      VirtualCells)

K_RANGE = 8  # range_check.rs:9


def _inv(x: int) -> int:
    return pow(x, -1, R_MOD)


# ---- utils.rs:47-108 -------------------------------------------------------------------------
def decompose_word_be(word: int, num_windows: int, window_bits: int) -> List[int]:
    mask = (1 << window_bits) - 1
    word &= (1 << (num_windows * window_bits)) - 1
    return [(word >> (window_bits * (num_windows - 1 - i))) & mask for i in range(num_windows)]


def to_u32(x: int) -> int:
    return x & 0xFFFFFFFF


# ---- halo2_gadgets LookupRangeCheckConfig<F, 8> + RangeCheckConfig (range_check.rs) -----------
class RangeCheck:
    def __init__(self, cs: ConstraintSystem, running_sum: Column, table_idx: Column):
        # LookupRangeCheckConfig::configure [UPSTREAM-RECALLED, halo2_gadgets v2023_04_20]
        cs.enable_equality(running_sum)
        self.q_lookup = cs.complex_selector()
        self.q_running = cs.complex_selector()
        self.q_bitshift = cs.selector()
        self.col = running_sum

        def lookup(m: VirtualCells):
            q_lookup = m.query_selector(self.q_lookup)
            q_running = m.query_selector(self.q_running)
            z_cur = m.query_advice(running_sum, 0)
            z_next = m.query_advice(running_sum, 1)
            running_sum_word = z_cur - z_next * (1 << K_RANGE)
            running_sum_lookup = q_running * running_sum_word
            q_short = Const(1) - q_running
            short_lookup = q_short * z_cur
            return [(q_lookup * (running_sum_lookup + short_lookup), table_idx)]
        cs.lookup("lookup", lookup)

        def bitshift(m: VirtualCells):
            q = m.query_selector(self.q_bitshift)
            word = m.query_advice(running_sum, -1)
            shifted = m.query_advice(running_sum, 0)
            inv_two_pow_s = m.query_advice(running_sum, 1)
            return q, [word * (1 << K_RANGE) * inv_two_pow_s - shifted]
        cs.create_gate("Short lookup bitshift", bitshift)

        # range_check.rs:34-51
        self.le_selector = cs.selector()

        def le(m: VirtualCells):
            q = m.query_selector(self.le_selector)
            x = m.query_advice(running_sum, -1)
            y = m.query_advice(running_sum, 0)
            diff = m.query_advice(running_sum, 1)
            return q, [x + diff - y]
        cs.create_gate("le", le)

    # halo2_gadgets copy_check -> range_check(words, strict)
    def copy_check(self, lay: SimpleFloorPlanner, element: Cell, num_words: int, strict: bool) -> List[Cell]:
        def body(r: Region):
            z0 = r.copy_advice(element, self.col, 0)
            zs = [z0]
            val = element.value
            inv = _inv(1 << K_RANGE)
            z = z0
            for idx in range(num_words):
                r.enable_selector(self.q_lookup, idx)
                r.enable_selector(self.q_running, idx)
                word = (val >> (K_RANGE * idx)) & ((1 << K_RANGE) - 1)
                zval = (z.value - word) * inv % R_MOD
                z = r.assign_advice(self.col, idx + 1, zval)
                zs.append(z)
            if strict:
                r.constants.append((0, zs[-1]))        # region.constrain_constant(zs.last(), 0)
            return zs
        return lay.assign_region("Range check", body)

    # halo2_gadgets copy_short_check -> short_range_check
    def copy_short_check(self, lay: SimpleFloorPlanner, element: Cell, num_bits: int):
        assert num_bits < K_RANGE

        def body(r: Region):
            el = r.copy_advice(element, self.col, 0)
            r.enable_selector(self.q_lookup, 0)        # lookup on the element itself
            r.enable_selector(self.q_lookup, 1)        # lookup on the shifted element
            r.enable_selector(self.q_bitshift, 1)
            shifted = el.value * (1 << (K_RANGE - num_bits)) % R_MOD
            r.assign_advice(self.col, 1, shifted)
            r.assign_advice_from_constant(self.col, 2, _inv(1 << num_bits))
        lay.assign_region("Range check %d bits" % num_bits, body)

    # range_check.rs:95-125
    def range_check(self, lay: SimpleFloorPlanner, cell: Cell, n_bits: int):
        words = n_bits // K_RANGE
        last = cell
        if words > 0:
            last = self.copy_check(lay, cell, words, n_bits % K_RANGE == 0)[-1]
        if n_bits % K_RANGE != 0:
            self.copy_short_check(lay, last, n_bits % K_RANGE)

    # range_check.rs:62-91
    def le_constant(self, lay: SimpleFloorPlanner, x: Cell, y: int):
        def body(r: Region):
            r.copy_advice(x, self.col, 0)
            r.assign_advice_from_constant(self.col, 1, y)
            diff = r.assign_advice(self.col, 2, (y - x.value) % R_MOD)
            r.enable_selector(self.le_selector, 1)
            return diff
        diff = lay.assign_region("le", body)
        self.range_check(lay, diff, y.bit_length())


def load_bytes_column(lay: SimpleFloorPlanner, table_column: Column):
    """range_check.rs:128-149: a table column holding 0..255 (only needed when no byte table exists already)."""
    lay.assign_table("table_idx", [table_column], [(i,) for i in range(1 << K_RANGE)])


# ---- greater_than.rs / encode_image.rs -------------------------------------------------------
class GreaterThan:
    def __init__(self, cs, x, y, diff, is_gt, rc: RangeCheck):
        self.x, self.y, self.diff, self.is_gt, self.rc = x, y, diff, is_gt, rc
        self.selector = cs.selector()                                   # greater_than.rs:80

        def gate(m):
            s = m.query_selector(self.selector)
            xx, yy = m.query_advice(x, 0), m.query_advice(y, 0)
            dd, gt = m.query_advice(diff, 0), m.query_advice(is_gt, 0)
            return s, [xx + dd - gt * Const(256) - yy]
        cs.create_gate("x + diff = 256 * is_gt + y", gate)

    def _gt(self, r: Region, x_cell: Cell, y: int):                     # greater_than.rs:106-131
        assert y <= 255
        gt = 1 if to_u32(x_cell.value) > y else 0
        diff = (256 * gt + y - x_cell.value) % R_MOD
        r.enable_selector(self.selector, 0)
        r.assign_advice_from_constant(self.y, 0, y)
        d = r.assign_advice(self.diff, 0, diff)
        g = r.assign_advice(self.is_gt, 0, gt)
        return d, g

    def witness(self, lay, x: int, y: int):                             # :135-165
        def body(r):
            xc = r.assign_advice(self.x, 0, x)
            d, g = self._gt(r, xc, y)
            return xc, d, g
        xc, d, g = lay.assign_region("greater_than_witness", body)
        self.rc.range_check(lay, xc, 8)
        self.rc.range_check(lay, g, 1)
        self.rc.range_check(lay, d, 8)
        return xc, g

    def copy(self, lay, x: Cell, y: int):                               # :167-191
        def body(r):
            xc = r.copy_advice(x, self.x, 0)
            return self._gt(r, xc, y)
        d, g = lay.assign_region("greater_than_copy", body)
        self.rc.range_check(lay, g, 1)
        self.rc.range_check(lay, d, 8)
        return g


class EncodeImage:
    def __init__(self, cs, x, y, diff, is_gt, rc, thresholds):
        self.gt = GreaterThan(cs, x, y, diff, is_gt, rc)
        self.col = is_gt
        self.thresholds = thresholds                                     # (w, h, b) ints in [0, 256]

    def encode(self, lay, image) -> List[Cell]:                          # encode_image.rs:75-150
        w, h, nb = self.thresholds.shape
        first = {}
        bits = []
        for b in range(nb):
            for i in range(w):
                for j in range(h):
                    t = int(self.thresholds[i, j, b])
                    assert t <= 256
                    if t == 0:
                        cell = lay.assign_region("bit is one", lambda r: r.assign_advice_from_constant(self.col, 0, 1))
                    elif (i, j) not in first:
                        xc, cell = self.gt.witness(lay, int(image[i, j]), t - 1)
                        first[(i, j)] = xc
                    else:
                        cell = self.gt.copy(lay, first[(i, j)], t - 1)
                    bits.append(cell)
        return bits


# ---- bits2num.rs -----------------------------------------------------------------------------
class Bits2Num:
    def __init__(self, cs, inp, acc):
        self.inp, self.acc = inp, acc
        self.selector = cs.selector()

        def gate(m):
            bit = m.query_advice(inp, 0)
            prev = m.query_advice(acc, 0)
            cur = m.query_advice(acc, 1)
            s = m.query_selector(self.selector)
            return s, [cur - (prev * 2 + bit)]
        cs.create_gate("next_num_constraint", gate)

    def convert_le(self, lay, bits: List[Cell]) -> Cell:
        bits = list(reversed(bits))

        def body(r):
            val = 0
            cell = r.assign_advice_from_constant(self.acc, 0, 0)
            for i, b in enumerate(bits):
                r.enable_selector(self.selector, i)
                val = (val * 2 + b.value) % R_MOD
                cell = r.assign_advice(self.acc, i + 1, val)
                r.copy_advice(b, self.inp, i)
            return cell
        return lay.assign_region("bits2num", body)


# ---- hash.rs ---------------------------------------------------------------------------------
class Hash:
    def __init__(self, cs, inp, quotient, remainder, msb, hsh, rc: RangeCheck, p: int, l: int, n_bits: int):
        self.cols = (inp, quotient, remainder, msb, hsh)
        self.rc, self.p, self.l, self.n_bits = rc, p, l, n_bits
        self.selector = cs.selector()

        def gate(m):
            s = m.query_selector(self.selector)
            i, q = m.query_advice(inp, 0), m.query_advice(quotient, 0)
            rem, ms, hh = m.query_advice(remainder, 0), m.query_advice(msb, 0), m.query_advice(hsh, 0)
            cubed = i * i * i
            mod_p = q * Const(p) + rem
            mod_2l = ms * Const(1 << l) + hh
            return s, [cubed - mod_p, rem - mod_2l]
        cs.create_gate("hash", gate)

    def hash(self, lay, inp: Cell) -> Cell:                              # hash.rs:129-210
        c_in, c_q, c_r, c_m, c_h = self.cols

        def body(r):
            r.enable_selector(self.selector, 0)
            ic = r.copy_advice(inp, c_in, 0)
            cubed = ic.value ** 3 % R_MOD
            q = cubed // self.p                                          # integer_division (utils.rs:47-58)
            rem = (cubed - q * self.p) % R_MOD
            msb = rem // (1 << self.l)
            hv = (rem - msb * (1 << self.l)) % R_MOD
            return (r.assign_advice(c_q, 0, q), r.assign_advice(c_r, 0, rem),
                    r.assign_advice(c_m, 0, msb), r.assign_advice(c_h, 0, hv))
        q, rem, msb, out = lay.assign_region("hash", body)
        self.rc.range_check(lay, q, self.n_bits * 3 - self.l)
        self.rc.range_check(lay, msb, 1)
        self.rc.le_constant(lay, rem, self.p - 1)
        return out


# ---- bloom_filter/array_lookup.rs --------------------------------------------------------------
class ArrayLookup:
    def __init__(self, cs, hash_dec, byte_index, bit_index, bloom_index, bloom_value, n_hashes, bits_per_hash,
                 word_index_bits=None):
        assert 7 <= bits_per_hash <= 32
        if word_index_bits is None:                     # ArrayLookupConfig::from(BloomFilterConfig), :51-75
            byte_index_bits = int((bits_per_hash - 3.0) / 2.0 - math.floor(math.log2(n_hashes)))   # :63-67
            word_index_bits = bits_per_hash - (byte_index_bits + 3)
        self.word_index_bits = word_index_bits
        self.n_hashes, self.bits_per_hash = n_hashes, bits_per_hash
        self.cols = (hash_dec, byte_index, bit_index, bloom_index, bloom_value)
        self.t_index, self.t_word, self.t_value = (cs.lookup_table_column() for _ in range(3))
        self.selector = cs.complex_selector()
        wib = self.word_index_bits

        def lookup(m):
            s = m.query_selector(self.selector)
            cur, nxt = m.query_advice(hash_dec, 0), m.query_advice(hash_dec, 1)
            byi, bii = m.query_advice(byte_index, 0), m.query_advice(bit_index, 0)
            current_hash = cur - nxt * (1 << bits_per_hash)
            word_index = (current_hash - byi * 8 - bii) * _inv(1 << (bits_per_hash - wib))
            bidx, bval = m.query_advice(bloom_index, 0), m.query_advice(bloom_value, 0)
            default, one = Const(R_MOD - 1), Const(1)

            def wd(x):
                return s * x + (one - s) * default
            return [(wd(bidx), self.t_index), (wd(word_index), self.t_word), (wd(bval), self.t_value)]
        cs.lookup("bloom filter lookup", lookup)
        self.words = None

    def bytes_per_word(self) -> int:
        return 1 << (self.bits_per_hash - self.word_index_bits - 3)

    def set_arrays(self, arrays):
        """arrays: (C*N, 2^bits_per_hash) bool -> big-endian packed words (from_be_bits)."""
        import numpy as np
        wl = 1 << (self.bits_per_hash - self.word_index_bits)        # bits per word (multiple of 8)
        a = np.asarray(arrays, dtype=np.uint8).reshape(arrays.shape[0], -1, wl)
        packed = np.packbits(a, axis=2, bitorder="big")               # first bit = most significant
        nbytes = wl // 8
        self.words = [[int.from_bytes(packed[i, j].tobytes(), "big") for j in range(packed.shape[1])]
                      for i in range(packed.shape[0])]
        assert packed.shape[2] == nbytes

    def load(self, lay):                                                  # :257-301
        rows = []
        for bi, ws in enumerate(self.words):
            for i, w in enumerate(ws):
                rows.append((bi, i, w))
        rows.append((R_MOD - 1, R_MOD - 1, R_MOD - 1))
        lay.assign_table("bloom_filters", [self.t_index, self.t_word, self.t_value], rows)

    def lookup(self, lay, hash_value: Cell, bloom_index: int):            # :305-456
        n_h, bph, wib = self.n_hashes, self.bits_per_hash, self.word_index_bits
        c_dec, c_byte, c_bit, c_bidx, c_val = self.cols

        def body(r):
            hashes_le = list(reversed(decompose_word_be(hash_value.value, n_h, bph)))
            idx = []
            for hv in hashes_le:
                hv = to_u32(hv)
                nbb = bph - wib
                idx.append((hv >> nbb, (hv & ((1 << nbb) - 1)) >> 3, hv & 7))
            values = [self.words[bloom_index][wi] for wi, _, _ in idx]
            dec = [hash_value.value]
            shift = _inv(1 << bph)
            for hv in hashes_le:
                dec.append((dec[-1] - hv) * shift % R_MOD)
            assert dec[-1] == 0
            for i, v in enumerate(dec):
                if i == 0:
                    r.copy_advice(hash_value, c_dec, 0)
                elif i < n_h:
                    r.assign_advice(c_dec, i, v)
                else:
                    r.assign_advice_from_constant(c_dec, i, 0)
            for i in range(n_h):
                r.assign_advice_from_constant(c_bidx, i, bloom_index)
            vcells = [r.assign_advice(c_val, i, v) for i, v in enumerate(values)]
            bycells, bicells = [], []
            for i, (_, by, bi) in enumerate(idx):
                bycells.append(r.assign_advice(c_byte, i, by))
                bicells.append(r.assign_advice(c_bit, i, bi))
            for i in range(n_h):
                r.enable_selector(self.selector, i)
            return list(reversed(list(zip(vcells, bycells, bicells))))
        return lay.assign_region("look up hash values", body)


# ---- bloom_filter/bit_selector.rs --------------------------------------------------------------
class BitSelector:
    def __init__(self, cs, byte, index, bit):
        self.cols = (byte, index, bit)
        self.selector = cs.complex_selector()
        self.byte_column, self.index_column, self.bit_column = (cs.lookup_table_column() for _ in range(3))

        def lookup(m):
            s = m.query_selector(self.selector)
            return [(s * m.query_advice(byte, 0), self.byte_column), (s * m.query_advice(index, 0), self.index_column),
                    (s * m.query_advice(bit, 0), self.bit_column)]
        cs.lookup("bit_lookup", lookup)

    def load(self, lay):                                                  # :57-95
        rows = [(b, i, 0 if b & (1 << (7 - i)) == 0 else 1) for b in range(256) for i in range(8)]
        lay.assign_table("byte,index,bit", [self.byte_column, self.index_column, self.bit_column], rows)

    def select(self, lay, byte: Cell, index: Cell) -> Cell:               # :138-164
        cb, ci, cbit = self.cols

        def body(r):
            bit = (byte.value >> (7 - to_u32(index.value))) & 1
            r.enable_selector(self.selector, 0)
            r.copy_advice(byte, cb, 0)
            r.copy_advice(index, ci, 0)
            return r.assign_advice(cbit, 0, bit)
        return lay.assign_region("select_bit", body)


# ---- bloom_filter/byte_selector.rs -------------------------------------------------------------
class ByteSelector:
    def __init__(self, cs, byte_dec, lookup_index, byte_index, byte_selector, selector_acc, byte_acc, byte_table):
        self.cols = (byte_dec, lookup_index, byte_index, byte_selector, selector_acc, byte_acc)
        self.s_dec = cs.complex_selector()
        self.s_bit, self.s_acc, self.s_right, self.s_byteacc = (cs.selector() for _ in range(4))

        def rec_byte(m):
            return m.query_advice(byte_dec, 0) - m.query_advice(byte_dec, 1) * (1 << 8)

        cs.lookup("byte_decomposition", lambda m: [(m.query_selector(self.s_dec) * rec_byte(m), byte_table)])

        def g_bit(m):
            s = m.query_selector(self.s_bit)
            b = m.query_advice(byte_selector, 0)
            return s, [b * b - b]
        cs.create_gate("selector_is_bit", g_bit)

        def g_acc(m):
            s = m.query_selector(self.s_acc)
            b = m.query_advice(byte_selector, 0)
            cur, nxt = m.query_advice(selector_acc, 0), m.query_advice(selector_acc, 1)
            return s, [nxt - cur - b]
        cs.create_gate("selector_acc", g_acc)

        def g_right(m):
            s = m.query_selector(self.s_right)
            li, bi, b = m.query_advice(lookup_index, 0), m.query_advice(byte_index, 0), m.query_advice(byte_selector, 0)
            return s, [b * (li - bi)]
        cs.create_gate("right_byte_selected", g_right)

        def g_byteacc(m):
            s = m.query_selector(self.s_byteacc)
            cur, nxt = m.query_advice(byte_acc, 0), m.query_advice(byte_acc, 1)
            byte = rec_byte(m)
            b = m.query_advice(byte_selector, 0)
            return s, [nxt - cur - b * byte]
        cs.create_gate("byte_acc", g_byteacc)

    def select(self, lay, word: Cell, index: Cell, num_bytes: int) -> Cell:   # :184-352
        c_dec, c_li, c_bi, c_sel, c_sacc, c_bacc = self.cols

        def body(r):
            bytes_be = decompose_word_be(word.value, num_bytes, 8)
            idx = to_u32(index.value)
            ith = bytes_be[idx]
            dec = [word.value]
            shift = _inv(1 << 8)
            for b in reversed(bytes_be):
                dec.append((dec[-1] - b) * shift % R_MOD)
            assert dec[-1] == 0
            for i, v in enumerate(dec):
                if i == 0:
                    r.copy_advice(word, c_dec, 0)
                elif i < num_bytes:
                    r.assign_advice(c_dec, i, v)
                else:
                    r.assign_advice_from_constant(c_dec, i, 0)
            for i in range(num_bytes):
                r.copy_advice(index, c_li, i)
            for i in range(num_bytes):
                r.assign_advice_from_constant(c_bi, num_bytes - 1 - i, i)
            for i in range(num_bytes):
                r.assign_advice(c_sel, i, 1 if (num_bytes - 1 - i) == idx else 0)
            for i in range(num_bytes + 1):
                if i == 0:
                    r.assign_advice_from_constant(c_sacc, 0, 0)
                elif i < num_bytes:
                    r.assign_advice(c_sacc, i, 1 if (num_bytes - i) <= idx else 0)
                else:
                    r.assign_advice_from_constant(c_sacc, i, 1)
            result = r.assign_advice_from_constant(c_bacc, 0, 0)
            for i in range(1, num_bytes + 1):
                result = r.assign_advice(c_bacc, i, ith if (num_bytes - i) <= idx else 0)
            for s in (self.s_dec, self.s_bit, self.s_acc, self.s_right, self.s_byteacc):
                for i in range(num_bytes):
                    r.enable_selector(s, i)
            return result
        return lay.assign_region("select_byte", body)


# ---- bloom_filter/and_bits.rs ------------------------------------------------------------------
class AndBits:
    def __init__(self, cs, bits, acc):
        self.bits, self.acc = bits, acc
        self.selector = cs.selector()

        def gate(m):
            s = m.query_selector(self.selector)
            bit, cur, nxt = m.query_advice(bits, 0), m.query_advice(acc, 0), m.query_advice(acc, 1)
            return s, [cur * bit - nxt]
        cs.create_gate("validate_bit_acc", gate)

    def and_bits(self, lay, bits: List[Cell]) -> Cell:                      # :83-121
        def body(r):
            accs = [1]
            for b in bits:
                accs.append(accs[-1] * b.value % R_MOD)
            for i, b in enumerate(bits):
                r.copy_advice(b, self.bits, i)
            cell = r.assign_advice_from_constant(self.acc, 0, 1)
            for i in range(1, len(accs)):
                cell = r.assign_advice(self.acc, i, accs[i])
                r.enable_selector(self.selector, i - 1)
            return cell
        return lay.assign_region("and bits", body)


# ---- bloom_filter.rs ---------------------------------------------------------------------------
class BloomFilter:
    def __init__(self, cs, adv: List[Column], n_hashes: int, bits_per_hash: int):      # :118-161
        self.array = ArrayLookup(cs, adv[0], adv[1], adv[2], adv[3], adv[4], n_hashes, bits_per_hash)
        self.bit = BitSelector(cs, adv[0], adv[1], adv[2])
        self.byte_column = self.bit.byte_column
        self.byte = ByteSelector(cs, adv[0], adv[1], adv[2], adv[3], adv[4], adv[5], self.byte_column)
        self.and_ = AndBits(cs, adv[4], adv[5])

    def load(self, lay):                                                     # :107-116
        self.array.load(lay)
        self.bit.load(lay)

    def bloom_lookup(self, lay, hash_value: Cell, bloom_index: int) -> Cell:  # :165-191
        bits = []
        for word, byte_index, bit_index in self.array.lookup(lay, hash_value, bloom_index):
            byte = self.byte.select(lay, word, byte_index, self.array.bytes_per_word())
            bits.append(self.bit.select(lay, byte, bit_index))
        return self.and_.and_bits(lay, bits)


# ---- response_accumulator.rs -------------------------------------------------------------------
class ResponseAccumulator:
    def __init__(self, cs, adv: List[Column]):
        self.adv = adv
        self.selector = cs.selector()

        def gate(m):
            s = m.query_selector(self.selector)
            x = [m.query_advice(adv[i], 0) for i in range(4)]
            prev, acc = m.query_advice(adv[4], 0), m.query_advice(adv[4], 1)
            return s, [x[0] + x[1] + x[2] + x[3] + prev - acc]
        cs.create_gate("accumulate_responses", gate)

    def accumulate(self, lay, responses: List[Cell]) -> Cell:                 # :77-133
        def body(r):
            cell = r.assign_advice_from_constant(self.adv[4], 0, 0)
            acc = 0
            nrows = (len(responses) + 3) // 4
            for row in range(nrows):
                r.enable_selector(self.selector, row)
                for i in range(4):
                    k = row * 4 + i
                    if k < len(responses):
                        r.copy_advice(responses[k], self.adv[i], row)
                        acc = (acc + responses[k].value) % R_MOD
                    else:
                        r.assign_advice_from_constant(self.adv[i], row, 0)
                cell = r.assign_advice(self.adv[4], row + 1, acc)
            return cell
        return lay.assign_region("accumulate_responses", body)


# ---- wnn.rs (gadget) ---------------------------------------------------------------------------
class WnnCircuit:
    """WnnCircuit::{configure_with_params, synthesize} + WnnChip (src/gadgets/wnn.rs:125-237, 334-393)."""

    def __init__(self, params: dict, bloom_filters, thresholds, input_permutation):
        self.params = params
        self.bloom_filters, self.thresholds, self.input_permutation = bloom_filters, thresholds, input_permutation
        cs = self.cs = ConstraintSystem()
        self.instance = cs.instance_column()
        adv = self.adv = [cs.advice_column() for _ in range(6)]
        for a in adv:
            cs.enable_equality(a)
        cs.enable_equality(self.instance)
        cs.enable_constant(cs.fixed_column())
        # WnnChip::configure (:125-172)
        self.bloom = BloomFilter(cs, adv, params["n_hashes"], params["bits_per_hash"])
        self.rc = RangeCheck(cs, adv[5], self.bloom.byte_column)
        self.encode = EncodeImage(cs, adv[0], adv[1], adv[2], adv[3], self.rc, thresholds)
        self.hash = Hash(cs, adv[0], adv[1], adv[2], adv[3], adv[4], self.rc, params["p"], params["l"], params["bits_per_filter"])
        self.accum = ResponseAccumulator(cs, adv[0:5])
        self.b2n = Bits2Num(cs, adv[3], adv[4])
        c, n, e = bloom_filters.shape
        self.n_classes, self.n_inputs = c, n
        self.bloom.array.set_arrays(bloom_filters.reshape(c * n, e))

    def synthesize(self, lay: SimpleFloorPlanner, image):
        self.bloom.load(lay)
        bit_cells = self.encode.encode(lay, image)
        permuted = [bit_cells[int(i)] for i in self.input_permutation]
        nb = self.params["bits_per_filter"]
        joint = [self.b2n.convert_le(lay, permuted[i:i + nb]) for i in range(0, len(permuted) - nb + 1, nb)]
        assert len(joint) == self.n_inputs
        hashes = [self.hash.hash(lay, j) for j in joint]
        scores = []
        for c in range(self.n_classes):
            resp = [self.bloom.bloom_lookup(lay, h, c * len(hashes) + i) for i, h in enumerate(hashes)]
            scores.append(resp)
        results = [self.accum.accumulate(lay, r) for r in scores]
        for i, cell in enumerate(results):
            lay.constrain_instance(cell, self.instance, i)
        return results
