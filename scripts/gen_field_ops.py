#!/usr/bin/env python3
"""Generator + bit-exact model of the straight-line PTX bodies in csrc/field_gen.cuh.

Two Montgomery primitives of the BN254 fields that the hand-written even/odd multiplier (field.cuh::fp_mul_eo) cannot
express, both built from the same carry-chain idiom (mad.lo.cc / madc.hi.cc pairs that ptxas fuses into
IMAD.WIDE.U32.X):

  * sqr  : a^2 / R           -- 36 wide products for the square (28 off-diagonal, doubled by a funnel shift, + 8
                                diagonal) and 64 for the reduction, against 128 for fp_mul_eo(a, a);
  * mul2 : (a*b + c*d) / R   -- two product rows per reduction row ("lazy reduction"): 192 wide products against
                                256 for two fp_mul_eo calls.  The result is < 3p (two conditional subtractions).

Each body is ONE asm statement (the carry flag never leaves it).  The same instruction list is (1) printed as PTX and
(2) interpreted here with exact carry-flag semantics, so the text that ptxas compiles is the text that the CPU tests
(tests/test_field_gen.py) check against Python big integers -- there is no GPU in the build container.

A third body, `mulk` (one level of subtractive Karatsuba on the 8 x 8-limb product: 48 + 64 wide products), is kept in
the generator and in the CPU tests but NOT in the shipped header: compiled for sm_100a (`--with-karatsuba`) it is 99
IMAD.WIDE + 13 IMAD.HI + 32 IMAD + ~160 ALU-pipe instructions per product = 512 FMA-pipe cycles against 544 for
fp_mul_eo (120 + 8 + 16 and ~58) -- 6 % fewer heavy-pipe cycles for almost three times the ALU-pipe work.

Usage:  python scripts/gen_field_ops.py            # rewrite 0g-halo2_b200/csrc/field_gen.cuh
        python scripts/gen_field_ops.py --check    # exit 1 if the committed header differs from the generator
"""
import os
import sys

M32 = 0xFFFFFFFF
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FIELDS = {
    "FrParams": R_MOD,
    "FqParams": Q_MOD,
}


def limbs(x, n=8):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def unlimbs(v):
    return sum(int(x) << (32 * i) for i, x in enumerate(v))


def mont_inv32(p):
    """-p^-1 mod 2^32"""
    return (-pow(p, -1, 1 << 32)) & M32


class Prog:
    """A straight-line PTX program over 32-bit virtual registers.

    Operands are strings (register names: inputs, outputs, temporaries) or ints (immediates)."""

    def __init__(self, inputs, outputs):
        self.inputs = list(inputs)
        self.outputs = list(outputs)
        self.ins = []
        self.ntmp = 0

    def tmp(self):
        self.ntmp += 1
        return "t%d" % (self.ntmp - 1)

    def op(self, name, dst, *src):
        self.ins.append((name, dst) + tuple(src))
        return dst

    # ---- interpretation -------------------------------------------------------------------------------
    def run(self, values):
        """values: {input name: int} -> {output name: int}.  Raises if the carry flag is read before it is written
        or a register is read before it is written."""
        reg = dict(values)
        cf = None

        def rd(x):
            if isinstance(x, int):
                return x & M32
            return reg[x]

        for ins in self.ins:
            name, dst, src = ins[0], ins[1], ins[2:]
            s = [rd(x) for x in src]
            base, *mods = name.split(".")
            carry_in = base.endswith("c") and base in ("addc", "subc", "madc")
            carry_out = "cc" in mods
            if carry_in:
                if cf is None:
                    raise RuntimeError("carry flag read before written: %r" % (ins,))
                cin = cf
            else:
                cin = 0
            if base in ("add", "addc"):
                t = s[0] + s[1] + cin
                res, co = t & M32, t >> 32
            elif base in ("sub", "subc"):
                t = s[0] - s[1] - cin
                res, co = t & M32, 1 if t < 0 else 0
            elif base in ("mad", "madc"):
                prod = s[0] * s[1]
                part = (prod & M32) if "lo" in mods else (prod >> 32)
                t = part + s[2] + cin
                res, co = t & M32, t >> 32
            elif base == "mul":
                prod = s[0] * s[1]
                res, co = ((prod & M32) if "lo" in mods else (prod >> 32)), None
            elif base == "shf":      # shf.l.wrap.b32 d, lo, hi, n  ->  upper word of (hi:lo) << n
                n = s[2] & 31
                res, co = ((((s[1] << 32) | s[0]) << n) >> 32) & M32, None
            elif base == "xor":
                res, co = s[0] ^ s[1], None
            elif base == "mov":
                res, co = s[0], None
            else:
                raise ValueError(name)
            assert co is None or co in (0, 1), ins
            reg[dst] = res
            if carry_out:
                cf = co
        return {o: reg[o] for o in self.outputs}

    # ---- PTX text -------------------------------------------------------------------------------------
    def ptx_lines(self):
        idx = {}
        for i, o in enumerate(self.outputs):
            idx[o] = "%%%d" % i
        for i, a in enumerate(self.inputs):
            idx[a] = "%%%d" % (len(self.outputs) + i)

        def fmt(x):
            if isinstance(x, int):
                return "0x%08x" % (x & M32)
            return idx.get(x, x)

        suffix = {"shf": ".b32", "xor": ".b32"}
        out = []
        for ins in self.ins:
            name, ops = ins[0], ins[1:]
            base = name.split(".")[0]
            out.append("%s%s %s;" % (name, suffix.get(base, ".u32"), ", ".join(fmt(x) for x in ops)))
        return out

    def count(self, prefix):
        return sum(1 for i in self.ins if i[0].startswith(prefix))


class Acc:
    """An array of limb accumulators whose entries are register names or None (= known zero)."""

    def __init__(self, n):
        self.v = [None] * n

    def get(self, i):
        return 0 if self.v[i] is None else self.v[i]


def wide_chain(p, arr, start, pairs, carry_in=False, tail=True):
    """One carry chain of wide multiply-adds: (arr[start+2k+1] : arr[start+2k]) += x_k * y_k for consecutive k,
    then the chain's carry into arr[start + 2 * len(pairs)] (when `tail` and that limb exists).
    `carry_in`: the chain continues the carry flag of the instruction emitted just before it."""
    fresh = all(arr.v[start + j] is None for j in range(2 * len(pairs))) and not carry_in
    for k, (x, y) in enumerate(pairs):
        lo, hi = start + 2 * k, start + 2 * k + 1
        dlo, dhi = p.tmp(), p.tmp()
        if fresh:
            p.op("mul.lo", dlo, x, y)
            p.op("mul.hi", dhi, x, y)
        else:
            first = k == 0 and not carry_in
            p.op("mad.lo.cc" if first else "madc.lo.cc", dlo, x, y, arr.get(lo))
            last = k == len(pairs) - 1
            end_no_carry = last and not (tail and start + 2 * len(pairs) < len(arr.v))
            p.op("madc.hi" if end_no_carry else "madc.hi.cc", dhi, x, y, arr.get(hi))
        arr.v[lo], arr.v[hi] = dlo, dhi
    e = start + 2 * len(pairs)
    if not fresh and tail and e < len(arr.v):
        d = p.tmp()
        p.op("addc", d, arr.get(e), 0)
        arr.v[e] = d


def reduce_row(p, u, w, pm, inv, carry_in):
    """One Montgomery row on T = U + 2^32 * W (field.cuh::eo_reduce): m = u0 * inv; T += m * p.  Afterwards u[0] == 0.
    `carry_in`: the carry of the `add.cc` that folded the left-over limb into u[0] feeds the first W chain."""
    m = p.tmp()
    p.op("mul.lo", m, u.get(0), inv)
    wide_chain(p, w, 0, [(m, pm[1]), (m, pm[3]), (m, pm[5]), (m, pm[7])], carry_in=carry_in, tail=False)
    # U chain: its carry (weight 2^256) lands in w[7]
    for k in range(4):
        dlo, dhi = p.tmp(), p.tmp()
        p.op("mad.lo.cc" if k == 0 else "madc.lo.cc", dlo, m, pm[2 * k], u.get(2 * k))
        p.op("madc.hi.cc", dhi, m, pm[2 * k], u.get(2 * k + 1))
        u.v[2 * k], u.v[2 * k + 1] = dlo, dhi
    d = p.tmp()
    p.op("addc", d, w.get(7), 0)
    w.v[7] = d


def shift_row(p, u, w, insert):
    """Divide T by 2^32 (u[0] == 0): new U = W, new W = U >> 64 with `insert` entering at relative limb 7, and the
    left-over limb u[1] added into the new u[0]; returns with the carry flag of that addition live."""
    x1 = u.get(1)
    nu, nw = Acc(8), Acc(8)
    nu.v = list(w.v)
    nw.v = list(u.v[2:8]) + [insert, None]
    d = p.tmp()
    p.op("add.cc", d, nu.get(0), x1)
    nu.v[0] = d
    return nu, nw


def final_combine(p, u, w, top, outs):
    """out = (U >> 32) + W (+ top at limb 7)"""
    for j in range(8):
        if j == 0:
            p.op("add.cc", outs[j], u.get(1), w.get(0))
        elif j < 7:
            p.op("addc.cc", outs[j], u.get(j + 1), w.get(j))
        else:
            p.op("addc", outs[j], w.get(7), 0 if top is None else top)


def gen_mul2(mod):
    """(a*b + c*d) / 2^256 mod p as a value < 3p: the even/odd CIOS of field.cuh with two product rows per reduction row.
    Invariant: T < 3p before every row, so T + a*b_i + c*d_i + m*p < 3p * 2^32 < 2^288 fits the 9-limb window."""
    pm, inv = limbs(mod), mont_inv32(mod)
    A = ["a%d" % i for i in range(8)]
    B = ["b%d" % i for i in range(8)]
    C = ["c%d" % i for i in range(8)]
    D = ["d%d" % i for i in range(8)]
    outs = ["r%d" % i for i in range(8)]
    p = Prog(A + B + C + D, outs)
    u, w = Acc(8), Acc(8)
    for i in range(8):
        carry = False
        if i > 0:
            u, w = shift_row(p, u, w, None)
            carry = True
        for (X, y) in ((A, B[i]), (C, D[i])):
            wide_chain(p, w, 0, [(X[1], y), (X[3], y), (X[5], y), (X[7], y)], carry_in=carry, tail=False)
            carry = False
            # U chain; carry out of limb 7 goes to w[7]
            fresh = all(x is None for x in u.v)
            for k in range(4):
                dlo, dhi = p.tmp(), p.tmp()
                if fresh:
                    p.op("mul.lo", dlo, X[2 * k], y)
                    p.op("mul.hi", dhi, X[2 * k], y)
                else:
                    p.op("mad.lo.cc" if k == 0 else "madc.lo.cc", dlo, X[2 * k], y, u.get(2 * k))
                    p.op("madc.hi.cc", dhi, X[2 * k], y, u.get(2 * k + 1))
                u.v[2 * k], u.v[2 * k + 1] = dlo, dhi
            if not fresh:
                t = p.tmp()
                p.op("addc", t, w.get(7), 0)
                w.v[7] = t
        reduce_row(p, u, w, pm, inv, carry_in=False)
    final_combine(p, u, w, None, outs)
    return p


def gen_sqr(mod):
    """a^2 / 2^256 mod p as a value < 2p.  Square first (16 limbs), then eight reduction rows on the low half while the
    high half enters the 9-limb window one limb per shift."""
    pm, inv = limbs(mod), mont_inv32(mod)
    A = ["a%d" % i for i in range(8)]
    outs = ["r%d" % i for i in range(8)]
    p = Prog(A, outs)
    # off-diagonal products a_i * a_j (i < j) at limb i + j: even positions in E, odd positions in O.  Chains are issued
    # in an order in which every chain's carry limb lies above everything written before, so a single `addc` suffices.
    E, O = Acc(16), Acc(16)
    for j in range(1, 8):
        ev = [(A[i], A[j]) for i in range(0, j, 2)]   # positions j, j+2, ...
        od = [(A[i], A[j]) for i in range(1, j, 2)]   # positions j+1, j+3, ...
        arr_ev, arr_od = (E, O) if j % 2 == 0 else (O, E)
        if ev:
            wide_chain(p, arr_ev, j, ev)
        if od:
            wide_chain(p, arr_od, j + 1, od)
    # S = E + O (limbs 1..15; limb 0 of both is zero)
    S = Acc(16)
    started = False
    for k in range(1, 16):
        e, o = E.v[k], O.v[k]
        if e is None and o is None and not started:
            continue
        if (e is None or o is None) and not started:
            S.v[k] = e if o is None else o
            continue
        d = p.tmp()
        if not started:
            p.op("add.cc", d, E.get(k), O.get(k))
            started = True
        elif k < 15:
            p.op("addc.cc", d, E.get(k), O.get(k))
        else:
            p.op("addc", d, E.get(k), O.get(k))
        S.v[k] = d
    # 2S by funnel shifts (S < 2^511: nothing leaves limb 15)
    S2 = Acc(16)
    for k in range(15, 0, -1):
        if S.v[k] is None and (k == 0 or S.v[k - 1] is None):
            continue
        d = p.tmp()
        p.op("shf.l.wrap", d, S.get(k - 1) if k >= 1 else 0, S.get(k), 1)
        S2.v[k] = d
    # t = 2S + sum a_i^2 * 2^(64 i): ONE carry chain of eight wide multiply-adds over all 16 limbs
    T = Acc(16)
    T.v = list(S2.v)
    for i in range(8):
        dlo, dhi = p.tmp(), p.tmp()
        if i == 0:
            p.op("mul.lo", dlo, A[0], A[0])           # limb 0 of 2S is zero
            p.op("mad.hi.cc", dhi, A[0], A[0], T.get(1))
        else:
            p.op("madc.lo.cc", dlo, A[i], A[i], T.get(2 * i))
            p.op("madc.hi.cc" if i < 7 else "madc.hi", dhi, A[i], A[i], T.get(2 * i + 1))
        T.v[2 * i], T.v[2 * i + 1] = dlo, dhi
    # reduction
    u, w = Acc(8), Acc(8)
    u.v = list(T.v[:8])
    for i in range(8):
        carry = False
        if i > 0:
            u, w = shift_row(p, u, w, T.v[7 + i])
            carry = True
        reduce_row(p, u, w, pm, inv, carry_in=carry)
    final_combine(p, u, w, T.v[15], outs)
    return p


def product4(p, X, Y):
    """4 x 4 limbs -> 8 limbs: even-position products in E, odd-position products in O (one carry chain per row and
    parity, issued so that every chain's carry limb lies above what was written before), then E + O."""
    E, O = Acc(8), Acc(8)
    for j in range(4):
        ev = [(X[0], Y[j]), (X[2], Y[j])]      # positions j, j + 2
        od = [(X[1], Y[j]), (X[3], Y[j])]      # positions j + 1, j + 3
        arr_ev, arr_od = (E, O) if j % 2 == 0 else (O, E)
        wide_chain(p, arr_ev, j, ev)
        wide_chain(p, arr_od, j + 1, od)
    Z = [E.v[0]]
    for k in range(1, 8):
        d = p.tmp()
        p.op("add.cc" if k == 1 else ("addc.cc" if k < 7 else "addc"), d, E.get(k), O.get(k))
        Z.append(d)
    return Z


def abs_diff4(p, X, Y):
    """|X - Y| over 4 limbs and the sign mask (0 or 0xffffffff)"""
    d = [p.tmp() for _ in range(4)]
    for k in range(4):
        p.op("sub.cc" if k == 0 else "subc.cc", d[k], X[k], Y[k])
    m = p.tmp()
    p.op("subc", m, 0, 0)
    x = [p.tmp() for _ in range(4)]
    for k in range(4):
        p.op("xor", x[k], d[k], m)
    r = [p.tmp() for _ in range(4)]
    for k in range(4):
        p.op("sub.cc" if k == 0 else ("subc.cc" if k < 3 else "subc"), r[k], x[k], m)
    return r, m


def gen_mulk(mod):
    """a*b / 2^256 mod p as a value < 2p with ONE level of (subtractive) Karatsuba on the 8 x 8-limb product: three 4 x 4
    products (48 wide multiplies instead of 64), then the reduction rows of gen_sqr."""
    pm, inv = limbs(mod), mont_inv32(mod)
    A = ["a%d" % i for i in range(8)]
    B = ["b%d" % i for i in range(8)]
    outs = ["r%d" % i for i in range(8)]
    p = Prog(A + B, outs)
    z0 = product4(p, A[:4], B[:4])
    z2 = product4(p, A[4:], B[4:])
    da, ma = abs_diff4(p, A[:4], A[4:])        # a_lo - a_hi
    db, mb = abs_diff4(p, B[4:], B[:4])        # b_hi - b_lo
    zm = product4(p, da, db)
    sg = p.tmp()
    p.op("xor", sg, ma, mb)                     # all ones when (a_lo - a_hi)(b_hi - b_lo) < 0
    # t = z0 + z2 (9 limbs)
    t = [p.tmp() for _ in range(9)]
    for k in range(8):
        p.op("add.cc" if k == 0 else "addc.cc", t[k], z0[k], z2[k])
    p.op("addc", t[8], 0, 0)
    # z1 = t +- zm = a_lo*b_hi + a_hi*b_lo (>= 0, < 2^257)
    zx = [p.tmp() for _ in range(8)]
    for k in range(8):
        p.op("xor", zx[k], zm[k], sg)
    cy = p.tmp()
    p.op("add.cc", cy, sg, 1)                   # carry flag := (sg != 0): the +1 of the two's complement
    z1 = [p.tmp() for _ in range(9)]
    for k in range(8):
        p.op("addc.cc", z1[k], t[k], zx[k])
    p.op("addc", z1[8], t[8], sg)
    # T = z0 + z1 * 2^128 + z2 * 2^256
    T = list(z0[:4])
    for k in range(4, 16):
        d = p.tmp()
        lo = z0[k] if k < 8 else z2[k - 8]
        hi = z1[k - 4] if k - 4 < 9 else 0
        p.op("add.cc" if k == 4 else ("addc.cc" if k < 15 else "addc"), d, lo, hi)
        T.append(d)
    u, w = Acc(8), Acc(8)
    u.v = list(T[:8])
    for i in range(8):
        carry = False
        if i > 0:
            u, w = shift_row(p, u, w, T[7 + i])
            carry = True
        reduce_row(p, u, w, pm, inv, carry_in=carry)
    final_combine(p, u, w, T[15], outs)
    return p


# ---- header --------------------------------------------------------------------------------------------
HEADER = '''// GENERATED by scripts/gen_field_ops.py -- do not edit; `python scripts/gen_field_ops.py --check` compares.
//
// Straight-line PTX bodies of two Montgomery primitives over the BN254 fields (see the generator for the algorithm and
// for the instruction-level model that tests/test_field_gen.py checks against Python big integers):
//   fp_sqr_gen<P>(a)          = a^2 / R mod p             (36 + 64 wide products instead of 128)
//   fp_mul2_gen<P>(a,b,c,d)   = (a*b + c*d) / R mod p     (one Montgomery reduction for two products: 192 instead of 256)
// Both return canonical values (< p), bit-identical to fp_mul / fp_add of field.cuh.
#pragma once
#if defined(__CUDA_ARCH__)
namespace zg {
template <class P> __device__ __forceinline__ Fp<P> fp_sqr_gen(const Fp<P>& a);
template <class P> __device__ __forceinline__ Fp<P> fp_mul2_gen(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d);
'''
HEADER_K = "template <class P> __device__ __forceinline__ Fp<P> fp_mulk_gen(const Fp<P>& a, const Fp<P>& b);\n"



def emit_function(kind, field, prog):
    L = []
    nout = len(prog.outputs)
    if kind == "sqr":
        L.append("template <> __device__ __forceinline__ Fp<%s> fp_sqr_gen<%s>(const Fp<%s>& a) {" % (field, field, field))
    elif kind == "mulk":
        L.append("template <> __device__ __forceinline__ Fp<%s> fp_mulk_gen<%s>(const Fp<%s>& a, const Fp<%s>& b) {" % ((field,) * 4))
    else:
        L.append("template <> __device__ __forceinline__ Fp<%s> fp_mul2_gen<%s>(const Fp<%s>& a, const Fp<%s>& b, "
                 "const Fp<%s>& c, const Fp<%s>& d) {" % ((field,) * 6))
    L.append("  uint32_t t[8];")
    L.append('  asm("{\\n\\t"')
    L.append('      ".reg .u32 t<%d>;\\n\\t"' % max(prog.ntmp, 1))
    for line in prog.ptx_lines():
        L.append('      "%s\\n\\t"' % line)
    L.append('      "}"')
    L.append("      : " + ", ".join('"=r"(t[%d])' % i for i in range(nout)))
    ins = []
    for name in prog.inputs:
        ins.append('"r"(%s.v[%s])' % (name[0], name[1:]))
    L.append("      : " + ", ".join(ins) + ");")
    L.append("  fp_final_sub<%s>(t);" % field)
    if kind == "mul2":
        L.append("  fp_final_sub<%s>(t);   // the two-product accumulator is < 3p" % field)
    L.append("  Fp<%s> r;" % field)
    L.append("#pragma unroll")
    L.append("  for (int i = 0; i < 8; i++) r.v[i] = t[i];")
    L.append("  return r;")
    L.append("}")
    return "\n".join(L)


def header_text():
    parts = [HEADER + (HEADER_K if "--with-karatsuba" in sys.argv else "")]
    for field, mod in FIELDS.items():
        for kind, gen in (("sqr", gen_sqr), ("mul2", gen_mul2)) + ((("mulk", gen_mulk),) if "--with-karatsuba" in sys.argv else ()):
            prog = gen(mod)
            parts.append("// %s %s: %d PTX instructions, %d multiply(-add) halves = %d wide products + 8 mul.lo for m" % (
                field, kind, len(prog.ins), prog.count("mad") + prog.count("mul"),
                (prog.count("mad") + prog.count("mul") - 8) // 2))
            parts.append(emit_function(kind, field, prog))
    parts.append("}  // namespace zg\n#endif\n")
    return "\n".join(parts)


def out_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "0g-halo2_b200", "csrc", "field_gen.cuh")


def main():
    text = header_text()
    path = out_path()
    if "--check" in sys.argv:
        cur = open(path).read() if os.path.exists(path) else ""
        if cur != text:
            print("field_gen.cuh differs from the generator output")
            return 1
        print("field_gen.cuh is up to date")
        return 0
    with open(path, "w") as f:
        f.write(text)
    print("wrote", os.path.normpath(path))
    return 0


if __name__ == "__main__":
    sys.exit(main())
