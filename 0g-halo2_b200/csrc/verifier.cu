// Host-side verifier: plonk::verify_proof with VerifierGWC / SingleStrategy over KZG<Bn256> and EvmTranscript, as
// instantiated by `Wnn::verify_proof` (/root/reference/src/wnn.rs:265-280; benched at benches/bench.rs:38-45).
//
// Restates halo2_proofs v2023_04_20 (un-vendored, /root/reference/Cargo.toml:21-25) `plonk/verifier.rs`,
// `poly/kzg/multiopen/gwc/verifier.rs`, `poly/kzg/strategy.rs` (DualMSM::check) and snark-verifier's EvmTranscript
// (read side).  Verification stays on the host in the reference as well (one small MSM and two pairings); nothing here
// launches a kernel, so a proof can be checked where no GPU exists.  The optimal-ate pairing below (tower
// Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - (9+u)), Fq12 = Fq6[w]/(w^2 - v); D-type twist; affine Miller loop over
// 6x+2, two Frobenius lines, final exponentiation = easy part by conjugation / inversion / p^2-Frobenius, hard part
// (p^4 - p^2 + 1)/r by the BN addition chain over f^x, f^(x^2), f^(x^3) and their Frobenius images) is
// written for this file; the test-only pairing of the CPU checker under oracle/ is the independent implementation
// (tests/test_verifier.py compares the two on bilinearity and on accept / reject of whole proofs).
#include <algorithm>
#include <array>
#include <memory>
#include <string>
#include <vector>
#include "curve.cuh"
#include "expr.cuh"
#include "../../include/zg_b200.h"

using namespace zg;

extern "C" void zg_debug_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);

namespace {

// ---- Fq2 / Fq6 / Fq12 -------------------------------------------------------------------------------------------
struct Fq2 { Fq c0, c1; };
struct Fq6 { Fq2 c0, c1, c2; };
struct Fq12 { Fq6 c0, c1; };

inline Fq fq0() { return fp_zero<FqParams>(); }
inline Fq fq1() { return fp_one<FqParams>(); }
inline Fq2 f2_zero() { return {fq0(), fq0()}; }
inline Fq2 f2_one() { return {fq1(), fq0()}; }
inline bool f2_is_zero(const Fq2& a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
inline bool f2_eq(const Fq2& a, const Fq2& b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
inline Fq2 f2_add(const Fq2& a, const Fq2& b) { return {fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
inline Fq2 f2_sub(const Fq2& a, const Fq2& b) { return {fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
inline Fq2 f2_neg(const Fq2& a) { return {fp_neg(a.c0), fp_neg(a.c1)}; }
inline Fq2 f2_conj(const Fq2& a) { return {a.c0, fp_neg(a.c1)}; }
inline Fq2 f2_dbl(const Fq2& a) { return {fp_dbl(a.c0), fp_dbl(a.c1)}; }
inline Fq2 f2_mul(const Fq2& a, const Fq2& b) {           // (a0 + a1 u)(b0 + b1 u), u^2 = -1, three products
  Fq t0 = fp_mul(a.c0, b.c0), t1 = fp_mul(a.c1, b.c1);
  Fq t2 = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
  return {fp_sub(t0, t1), fp_sub(fp_sub(t2, t0), t1)};
}
inline Fq2 f2_sqr(const Fq2& a) {                          // (a0 + a1)(a0 - a1) + 2 a0 a1 u
  Fq t = fp_mul(a.c0, a.c1);
  return {fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1)), fp_dbl(t)};
}
inline Fq2 f2_mul_fq(const Fq2& a, const Fq& s) { return {fp_mul(a.c0, s), fp_mul(a.c1, s)}; }
inline Fq2 f2_mul_xi(const Fq2& a) {                       // * (9 + u) = (9 a0 - a1) + (a0 + 9 a1) u
  Fq a0_8 = fp_dbl(fp_dbl(fp_dbl(a.c0))), a1_8 = fp_dbl(fp_dbl(fp_dbl(a.c1)));
  return {fp_sub(fp_add(a0_8, a.c0), a.c1), fp_add(fp_add(a1_8, a.c1), a.c0)};
}
inline Fq2 f2_inv(const Fq2& a) {                          // conj(a) / (a0^2 + a1^2)
  Fq d = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
  return {fp_mul(a.c0, d), fp_neg(fp_mul(a.c1, d))};
}

inline Fq6 f6_zero() { return {f2_zero(), f2_zero(), f2_zero()}; }
inline Fq6 f6_one() { return {f2_one(), f2_zero(), f2_zero()}; }
inline Fq6 f6_add(const Fq6& a, const Fq6& b) { return {f2_add(a.c0, b.c0), f2_add(a.c1, b.c1), f2_add(a.c2, b.c2)}; }
inline Fq6 f6_sub(const Fq6& a, const Fq6& b) { return {f2_sub(a.c0, b.c0), f2_sub(a.c1, b.c1), f2_sub(a.c2, b.c2)}; }
inline Fq6 f6_neg(const Fq6& a) { return {f2_neg(a.c0), f2_neg(a.c1), f2_neg(a.c2)}; }
inline Fq6 f6_mul_v(const Fq6& a) { return {f2_mul_xi(a.c2), a.c0, a.c1}; }   // v^3 = xi
Fq6 f6_mul(const Fq6& a, const Fq6& b) {
  Fq2 a0b0 = f2_mul(a.c0, b.c0), a1b1 = f2_mul(a.c1, b.c1), a2b2 = f2_mul(a.c2, b.c2);
  // cross terms by Karatsuba: a1b2 + a2b1 = (a1+a2)(b1+b2) - a1b1 - a2b2, ...
  Fq2 x12 = f2_sub(f2_sub(f2_mul(f2_add(a.c1, a.c2), f2_add(b.c1, b.c2)), a1b1), a2b2);
  Fq2 x01 = f2_sub(f2_sub(f2_mul(f2_add(a.c0, a.c1), f2_add(b.c0, b.c1)), a0b0), a1b1);
  Fq2 x02 = f2_sub(f2_sub(f2_mul(f2_add(a.c0, a.c2), f2_add(b.c0, b.c2)), a0b0), a2b2);
  return {f2_add(a0b0, f2_mul_xi(x12)), f2_add(x01, f2_mul_xi(a2b2)), f2_add(x02, a1b1)};
}
Fq6 f6_inv(const Fq6& a) {
  Fq2 t0 = f2_sub(f2_sqr(a.c0), f2_mul_xi(f2_mul(a.c1, a.c2)));
  Fq2 t1 = f2_sub(f2_mul_xi(f2_sqr(a.c2)), f2_mul(a.c0, a.c1));
  Fq2 t2 = f2_sub(f2_sqr(a.c1), f2_mul(a.c0, a.c2));
  Fq2 d = f2_add(f2_mul(a.c0, t0), f2_mul_xi(f2_add(f2_mul(a.c2, t1), f2_mul(a.c1, t2))));
  Fq2 di = f2_inv(d);
  return {f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di)};
}

inline Fq12 f12_one() { return {f6_one(), f6_zero()}; }
Fq12 f12_mul(const Fq12& a, const Fq12& b) {               // w^2 = v
  Fq6 t0 = f6_mul(a.c0, b.c0), t1 = f6_mul(a.c1, b.c1);
  Fq6 t2 = f6_mul(f6_add(a.c0, a.c1), f6_add(b.c0, b.c1));
  return {f6_add(t0, f6_mul_v(t1)), f6_sub(f6_sub(t2, t0), t1)};
}
inline Fq12 f12_sqr(const Fq12& a) {                       // (c0 + c1 w)^2 = c0^2 + v c1^2 + 2 c0 c1 w  (complex squaring)
  Fq6 ab = f6_mul(a.c0, a.c1);
  Fq6 s = f6_mul(f6_add(a.c0, a.c1), f6_add(a.c0, f6_mul_v(a.c1)));
  return {f6_sub(f6_sub(s, ab), f6_mul_v(ab)), f6_add(ab, ab)};
}
inline Fq12 f12_conj(const Fq12& a) { return {a.c0, f6_neg(a.c1)}; }
Fq12 f12_inv(const Fq12& a) {
  Fq6 d = f6_inv(f6_sub(f6_mul(a.c0, a.c0), f6_mul_v(f6_mul(a.c1, a.c1))));
  return {f6_mul(a.c0, d), f6_neg(f6_mul(a.c1, d))};
}
bool f12_is_one(const Fq12& a) {
  const Fq12 o = f12_one();
  return f2_eq(a.c0.c0, o.c0.c0) && f2_is_zero(a.c0.c1) && f2_is_zero(a.c0.c2) && f2_is_zero(a.c1.c0) && f2_is_zero(a.c1.c1) &&
         f2_is_zero(a.c1.c2);
}

// ---- big integers as hex strings (constants of the curve; parsed once) -----------------------------------------------
std::vector<uint8_t> hex_bits_msb_first(const char* hex) {
  std::vector<uint8_t> bits;
  for (const char* p = hex; *p; p++) {
    int v = (*p >= '0' && *p <= '9') ? *p - '0' : (*p >= 'a' && *p <= 'f') ? *p - 'a' + 10 : *p - 'A' + 10;
    for (int b = 3; b >= 0; b--) bits.push_back((uint8_t)((v >> b) & 1));
  }
  size_t lead = 0;
  while (lead < bits.size() && !bits[lead]) lead++;
  bits.erase(bits.begin(), bits.begin() + lead);
  return bits;
}
Fq2 f2_pow_bits(const Fq2& a, const std::vector<uint8_t>& bits) {
  Fq2 r = f2_one();
  for (uint8_t b : bits) {
    r = f2_sqr(r);
    if (b) r = f2_mul(r, a);
  }
  return r;
}

// BN254: x = 4965661367192848881; p = 36x^4 + 36x^3 + 24x^2 + 6x + 1; r = 36x^4 + 36x^3 + 18x^2 + 6x + 1
const char* const HEX_6X_PLUS_2 = "19d797039be763ba8";
const char* const HEX_P_MINUS_1_OVER_3 = "10216f7ba065e00de81ac1e7808072c9dd2b2385cd7b438469602eb24829a9c2";
const char* const HEX_P_MINUS_1_OVER_2 = "183227397098d014dc2822db40c0ac2ecbc0b548b438e5469e10460b6c3e7ea3";
const char* const HEX_P_MINUS_1_OVER_6 = "810b7bdd032f006f40d60f3c0403964ee9591c2e6bda1c234b017592414d4e1";
const char* const HEX_X = "44e992b44a6909f1";   // the BN parameter x

struct G2Aff { Fq2 x, y; bool inf; };

struct PairingConsts {
  std::vector<uint8_t> loop_bits, x_bits;
  Fq2 frob_x, frob_y;       // xi^((p-1)/3), xi^((p-1)/2): the p-power Frobenius on twist coordinates
  Fq2 gamma[6];             // gamma[i] = xi^(i (p-1)/6): the p-power Frobenius on Fq12 = Fq2[W]/(W^6 - xi)
  Fq2 twist_b;              // 3 / (9 + u)
  PairingConsts() {
    loop_bits = hex_bits_msb_first(HEX_6X_PLUS_2);
    x_bits = hex_bits_msb_first(HEX_X);
    Fq nine = fp_from_u64<FqParams>(9);
    Fq2 xi{nine, fq1()};
    frob_x = f2_pow_bits(xi, hex_bits_msb_first(HEX_P_MINUS_1_OVER_3));
    frob_y = f2_pow_bits(xi, hex_bits_msb_first(HEX_P_MINUS_1_OVER_2));
    gamma[0] = f2_one();
    gamma[1] = f2_pow_bits(xi, hex_bits_msb_first(HEX_P_MINUS_1_OVER_6));
    for (int i = 2; i < 6; i++) gamma[i] = f2_mul(gamma[i - 1], gamma[1]);
    twist_b = f2_mul_fq(f2_inv(xi), fp_from_u64<FqParams>(3));
  }
};
const PairingConsts& consts() {
  static const PairingConsts c;
  return c;
}

bool g2_on_curve(const G2Aff& q) {
  if (q.inf) return true;
  return f2_eq(f2_sqr(q.y), f2_add(f2_mul(f2_sqr(q.x), q.x), consts().twist_b));
}
G2Aff g2_frobenius(const G2Aff& q) {
  if (q.inf) return q;
  return {f2_mul(f2_conj(q.x), consts().frob_x), f2_mul(f2_conj(q.y), consts().frob_y), false};
}
G2Aff g2_neg(const G2Aff& q) { return {q.x, f2_neg(q.y), q.inf}; }

// One Miller step on twist coordinates: the line through T and Q (Q == T: tangent) evaluated at P = (xP, yP) in G1,
//   l(P) = yP - lambda xP w + (lambda xT - yT) w^3        (w^3 = v w; sparse: c0.c0, c1.c0, c1.c1)
// and T <- T + Q.  Vertical lines / points at infinity multiply by 1 (their value lies in a proper subfield and is
// erased by the final exponentiation).
Fq12 line_and_add(G2Aff& T, const G2Aff& Q, const Fq& xP, const Fq& yP, bool dbl) {
  if (T.inf || Q.inf) {
    if (T.inf) T = Q;
    return f12_one();
  }
  Fq2 lambda;
  if (dbl || (f2_eq(T.x, Q.x) && f2_eq(T.y, Q.y))) {
    if (f2_is_zero(T.y)) { T.inf = true; return f12_one(); }
    Fq2 x2 = f2_sqr(T.x);
    lambda = f2_mul(f2_add(f2_dbl(x2), x2), f2_inv(f2_dbl(T.y)));
  } else {
    if (f2_eq(T.x, Q.x)) { T.inf = true; return f12_one(); }
    lambda = f2_mul(f2_sub(Q.y, T.y), f2_inv(f2_sub(Q.x, T.x)));
  }
  Fq12 l;
  l.c0 = f6_zero();
  l.c1 = f6_zero();
  l.c0.c0 = {yP, fq0()};
  l.c1.c0 = f2_neg(f2_mul_fq(lambda, xP));
  l.c1.c1 = f2_sub(f2_mul(lambda, T.x), T.y);
  Fq2 x3 = f2_sub(f2_sub(f2_sqr(lambda), T.x), Q.x);
  Fq2 y3 = f2_sub(f2_mul(lambda, f2_sub(T.x, x3)), T.y);
  T.x = x3;
  T.y = y3;
  return l;
}

// a^p for a = sum_i a_i W^i (a_i in Fq2; c0 = (a_0, a_2, a_4), c1 = (a_1, a_3, a_5)): conj(a_i) * gamma[i]
Fq12 f12_frobenius(const Fq12& a) {
  const PairingConsts& C = consts();
  Fq12 r;
  r.c0.c0 = f2_conj(a.c0.c0);
  r.c0.c1 = f2_mul(f2_conj(a.c0.c1), C.gamma[2]);
  r.c0.c2 = f2_mul(f2_conj(a.c0.c2), C.gamma[4]);
  r.c1.c0 = f2_mul(f2_conj(a.c1.c0), C.gamma[1]);
  r.c1.c1 = f2_mul(f2_conj(a.c1.c1), C.gamma[3]);
  r.c1.c2 = f2_mul(f2_conj(a.c1.c2), C.gamma[5]);
  return r;
}
Fq12 f12_pow_x(const Fq12& a) {
  Fq12 r = f12_one();
  for (uint8_t b : consts().x_bits) {
    r = f12_sqr(r);
    if (b) r = f12_mul(r, a);
  }
  return r;
}
// f^((p^12 - 1)/r).  Easy part: f^((p^6 - 1)(p^2 + 1)); afterwards the inverse is the conjugate.  Hard part
// (p^4 - p^2 + 1)/r = p + p^2 + p^3 - 2 + 6 x^2 p^2 - 12 x p - 18 (x + x^2 p) - 30 x^2 - 36 (x^3 + x^3 p)  (checked as an
// integer identity for this x), evaluated with the exponent vector (1, 2, 6, 12, 18, 30, 36) addition chain.
Fq12 final_exponentiation(const Fq12& f) {
  Fq12 g = f12_mul(f12_conj(f), f12_inv(f));                 // f^(p^6 - 1)
  g = f12_mul(f12_frobenius(f12_frobenius(g)), g);           // ^(p^2 + 1)
  const Fq12 gx = f12_pow_x(g), gx2 = f12_pow_x(gx), gx3 = f12_pow_x(gx2);
  const Fq12 gp = f12_frobenius(g), gp2 = f12_frobenius(gp), gp3 = f12_frobenius(gp2);
  const Fq12 y0 = f12_mul(f12_mul(gp, gp2), gp3);
  const Fq12 y1 = f12_conj(g);
  const Fq12 y2 = f12_frobenius(f12_frobenius(gx2));
  const Fq12 y3 = f12_conj(f12_frobenius(gx));
  const Fq12 y4 = f12_conj(f12_mul(gx, f12_frobenius(gx2)));
  const Fq12 y5 = f12_conj(gx2);
  const Fq12 y6 = f12_conj(f12_mul(gx3, f12_frobenius(gx3)));
  Fq12 t0 = f12_mul(f12_mul(f12_sqr(y6), y4), y5);
  Fq12 t1 = f12_mul(f12_mul(y3, y5), t0);
  t0 = f12_mul(t0, y2);
  t1 = f12_sqr(f12_mul(f12_sqr(t1), t0));
  t0 = f12_mul(t1, y1);
  t1 = f12_mul(t1, y0);
  return f12_mul(f12_sqr(t0), t1);
}

// prod_i e(P_i, Q_i) == 1 ?   (P_i in G1 affine, Q_i on the twist)
bool pairing_product_is_one(const std::vector<G1Affine>& P, const std::vector<G2Aff>& Q) {
  const PairingConsts& C = consts();
  const size_t n = P.size();
  std::vector<G2Aff> T(n);
  std::vector<char> live(n);
  for (size_t i = 0; i < n; i++) {
    live[i] = !(affine_is_identity(P[i]) || Q[i].inf);     // e(O, Q) = e(P, O) = 1
    T[i] = Q[i];
  }
  Fq12 f = f12_one();
  for (size_t b = 1; b < C.loop_bits.size(); b++) {
    f = f12_sqr(f);
    for (size_t i = 0; i < n; i++)
      if (live[i]) f = f12_mul(f, line_and_add(T[i], T[i], P[i].x, P[i].y, true));
    if (C.loop_bits[b])
      for (size_t i = 0; i < n; i++)
        if (live[i]) f = f12_mul(f, line_and_add(T[i], Q[i], P[i].x, P[i].y, false));
  }
  for (size_t i = 0; i < n; i++) {
    if (!live[i]) continue;
    G2Aff q1 = g2_frobenius(Q[i]);
    G2Aff q2 = g2_neg(g2_frobenius(q1));
    f = f12_mul(f, line_and_add(T[i], q1, P[i].x, P[i].y, false));
    f = f12_mul(f, line_and_add(T[i], q2, P[i].x, P[i].y, false));
  }
  return f12_is_one(final_exponentiation(f));
}

// ---- G1 helpers on the host ----------------------------------------------------------------------------------------------
bool g1_on_curve(const G1Affine& p) {
  return fp_eq(fp_sqr(p.y), fp_add(fp_mul(fp_sqr(p.x), p.x), fp_from_u64<FqParams>(3)));
}
G1Affine g1_to_affine(const G1Xyzz& p) {
  G1Affine a;
  if (xyzz_is_identity(p)) { a.x = fq0(); a.y = fq0(); return a; }
  Fq i3 = fp_inv(p.zzz);
  Fq t = fp_mul(p.zz, i3);                 // 1 / sqrt(ZZ) ... (ZZ * ZZZ^-1)^2 = 1 / ZZ
  a.x = fp_mul(p.x, fp_sqr(t));
  a.y = fp_mul(p.y, i3);
  return a;
}
// sum_i s_i P_i with one shared doubling chain (4-bit windows); scalars in Montgomery form
G1Xyzz g1_msm(const std::vector<G1Affine>& pts, const std::vector<Fr>& scalars) {
  const size_t n = pts.size();
  std::vector<std::array<G1Xyzz, 16>> tab(n);
  std::vector<Fr> can(n);
  for (size_t i = 0; i < n; i++) {
    can[i] = fp_from_mont(scalars[i]);
    tab[i][0] = xyzz_identity();
    G1Xyzz base = xyzz_from_affine(pts[i]);
    for (int j = 1; j < 16; j++) {
      tab[i][j] = tab[i][j - 1];
      xyzz_add(tab[i][j], base);
    }
  }
  G1Xyzz acc = xyzz_identity();
  for (int nib = 63; nib >= 0; nib--) {
    for (int d = 0; d < 4; d++) acc = xyzz_double(acc);
    for (size_t i = 0; i < n; i++) {
      const uint32_t d = (can[i].v[nib >> 3] >> (4 * (nib & 7))) & 15u;
      if (d) xyzz_add(acc, tab[i][d]);
    }
  }
  return acc;
}

// ---- Fr helpers --------------------------------------------------------------------------------------------------------------
inline Fr fr0() { return fp_zero<FrParams>(); }
inline Fr fr1() { return fp_one<FrParams>(); }
Fr fr_pow_u64(const Fr& a, uint64_t e) { return fp_pow_var(a, e); }
Fr fr_delta() {
  Fr raw;
  const uint32_t v[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
  for (int i = 0; i < 8; i++) raw.v[i] = v[i];
  return fp_to_mont(raw);
}
Fr fr_root_of_unity(uint32_t k) {          // halo2curves Fr::ROOT_OF_UNITY ^ (2^(28-k))
  Fr raw;
  const uint32_t v[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
  for (int i = 0; i < 8; i++) raw.v[i] = v[i];
  Fr w = fp_to_mont(raw);
  for (uint32_t i = k; i < 28; i++) w = fp_sqr(w);
  return w;
}

// ---- transcript (read side of snark-verifier's EvmTranscript<G1Affine, NativeLoader, _, _>) ---------------------------------------
template <class P>
bool from_be32(const uint8_t* b, Fp<P>& out) {    // canonical big-endian bytes -> Montgomery; false when >= modulus
  Fp<P> raw;
  for (int i = 0; i < 8; i++) {
    uint32_t w = 0;
    for (int j = 0; j < 4; j++) w |= (uint32_t)b[31 - (4 * i + j)] << (8 * j);
    raw.v[i] = w;
  }
  for (int i = 7; i >= 0; i--) {
    if (raw.v[i] < P::mod(i)) { out = fp_to_mont(raw); return true; }
    if (raw.v[i] > P::mod(i)) return false;
  }
  return false;
}
template <class P>
void to_be32(const Fp<P>& mont, uint8_t out[32]) {
  Fp<P> c = fp_from_mont(mont);
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) out[31 - (4 * i + b)] = (uint8_t)(c.v[i] >> (8 * b));
}
struct Reader {
  const uint8_t* p;
  size_t len, pos = 0;
  std::vector<uint8_t> buf;
  bool ok = true;
  void common_scalar(const Fr& s) {
    uint8_t b[32];
    to_be32(s, b);
    buf.insert(buf.end(), b, b + 32);
  }
  G1Affine read_point() {
    G1Affine a;
    a.x = fq0(); a.y = fq0();
    if (!ok || pos + 64 > len) { ok = false; return a; }
    if (!from_be32(p + pos, a.x) || !from_be32(p + pos + 32, a.y) || !g1_on_curve(a)) { ok = false; return a; }
    buf.insert(buf.end(), p + pos, p + pos + 64);
    pos += 64;
    return a;
  }
  Fr read_scalar() {
    Fr s = fr0();
    if (!ok || pos + 32 > len) { ok = false; return s; }
    if (!from_be32(p + pos, s)) { ok = false; return s; }
    buf.insert(buf.end(), p + pos, p + pos + 32);
    pos += 32;
    return s;
  }
  Fr squeeze() {
    std::vector<uint8_t> d(buf);
    if (buf.size() == 32) d.push_back(1);
    uint8_t h[32];
    zg_debug_keccak256(d.data(), d.size(), h);
    buf.assign(h, h + 32);
    Fr raw;
    for (int i = 0; i < 8; i++) {
      uint32_t w = 0;
      for (int b = 0; b < 4; b++) w |= (uint32_t)h[31 - (4 * i + b)] << (8 * b);
      raw.v[i] = w;
    }
    Fr r2;
    for (int i = 0; i < 8; i++) r2.v[i] = FrParams::r2(i);
    return fp_mul(r2, raw);              // (value mod r) in Montgomery form; the raw operand may exceed r
  }
};

}  // namespace

struct zg_vk {
  uint32_t k = 0, A = 0, F = 0, I = 0, degree = 0, bf = 0, m = 0, n_progs = 0, n_gate_progs = 0, n_lookups = 0;
  std::vector<std::pair<uint32_t, int32_t>> q[3];
  std::vector<std::pair<uint32_t, uint32_t>> perm;
  std::vector<uint32_t> prog_off, ops, in_first, in_count, tab_first, tab_count;
  std::vector<Fr> constants;
  std::vector<G1Affine> fixed_comm, perm_comm;
  Fr transcript_repr;
  std::string err;
  int fail(int code, const char* msg) { err = msg; return code; }
};

namespace {

int vk_parse(zg_vk* vk, const uint32_t* w, size_t nw) {
  size_t p = 0;
  auto need = [&](size_t c) { return p + c <= nw; };
  if (!need(6) || w[0] != 0x5A473031u) return vk->fail(ZG_E_INVALID, "vk: bad constraint-system blob");
  vk->A = w[1]; vk->F = w[2]; vk->I = w[3]; vk->degree = w[4]; vk->bf = w[5];
  p = 6;
  for (int kind = 0; kind < 3; kind++) {
    if (!need(1)) return vk->fail(ZG_E_INVALID, "vk: truncated blob");
    uint32_t c = w[p++];
    if (!need(2 * (size_t)c)) return vk->fail(ZG_E_INVALID, "vk: truncated blob");
    for (uint32_t i = 0; i < c; i++, p += 2) vk->q[kind].push_back({w[p], (int32_t)w[p + 1]});
  }
  if (!need(1)) return vk->fail(ZG_E_INVALID, "vk: truncated blob");
  vk->m = w[p++];
  if (!need(2 * (size_t)vk->m + 1)) return vk->fail(ZG_E_INVALID, "vk: truncated blob");
  for (uint32_t i = 0; i < vk->m; i++, p += 2) vk->perm.push_back({w[p], w[p + 1]});
  vk->n_progs = w[p++];
  if (!need((size_t)vk->n_progs + 3)) return vk->fail(ZG_E_INVALID, "vk: truncated blob");
  vk->prog_off.assign(w + p, w + p + vk->n_progs + 1);
  p += vk->n_progs + 1;
  vk->n_gate_progs = w[p++];
  vk->n_lookups = w[p++];
  if (!need(4 * (size_t)vk->n_lookups + 1)) return vk->fail(ZG_E_INVALID, "vk: truncated blob");
  for (uint32_t l = 0; l < vk->n_lookups; l++, p += 4) {
    vk->in_first.push_back(w[p]); vk->in_count.push_back(w[p + 1]);
    vk->tab_first.push_back(w[p + 2]); vk->tab_count.push_back(w[p + 3]);
  }
  uint32_t nops = w[p++];
  if (!need(nops)) return vk->fail(ZG_E_INVALID, "vk: truncated blob");
  vk->ops.assign(w + p, w + p + nops);
  if (vk->degree < 3 || vk->n_gate_progs > vk->n_progs) return vk->fail(ZG_E_INVALID, "vk: inconsistent blob");
  const uint32_t ncols[3] = {vk->A, vk->F, vk->I};
  for (int kind = 0; kind < 3; kind++)
    for (auto& qq : vk->q[kind])
      if (qq.first >= ncols[kind]) return vk->fail(ZG_E_INVALID, "vk: query names a column that does not exist");
  for (auto& pc : vk->perm)
    if (pc.first > 2 || pc.second >= ncols[pc.first]) return vk->fail(ZG_E_INVALID, "vk: bad permutation column");
  for (uint32_t g = 0; g < vk->n_progs; g++) {
    if (vk->prog_off[g] > vk->prog_off[g + 1] || vk->prog_off[g + 1] > nops) return vk->fail(ZG_E_INVALID, "vk: bad program table");
    int depth = 0;
    for (uint32_t pc = vk->prog_off[g]; pc < vk->prog_off[g + 1]; pc++) {
      const uint32_t op = vk->ops[pc] & 0xff, arg = vk->ops[pc] >> 8;
      if (op == OP_CONST || op == OP_SCALE) {
        if (arg >= vk->constants.size()) return vk->fail(ZG_E_INVALID, "vk: constant index out of range");
      } else if (op >= OP_ADVICE && op <= OP_INSTANCE) {
        if (arg >= vk->q[op - OP_ADVICE].size()) return vk->fail(ZG_E_INVALID, "vk: query index out of range");
      } else if (op != OP_NEG && op != OP_ADD && op != OP_SUB && op != OP_MUL) {
        return vk->fail(ZG_E_INVALID, "vk: unknown opcode");
      }
      if (op <= OP_INSTANCE) depth++;
      else if (op == OP_ADD || op == OP_SUB || op == OP_MUL) depth--;
      if (depth < 1) return vk->fail(ZG_E_INVALID, "vk: malformed program");
    }
    if (depth != 1) return vk->fail(ZG_E_INVALID, "vk: malformed program");
  }
  for (uint32_t l = 0; l < vk->n_lookups; l++)
    if ((uint64_t)vk->in_first[l] + vk->in_count[l] > vk->n_progs || (uint64_t)vk->tab_first[l] + vk->tab_count[l] > vk->n_progs ||
        !vk->in_count[l] || !vk->tab_count[l])
      return vk->fail(ZG_E_INVALID, "vk: lookup program range out of bounds");
  return ZG_OK;
}

// one RPN program at the opening point: query values come from the proof's evaluations
Fr eval_at_point(const zg_vk* vk, uint32_t prog, const std::vector<Fr> ev[3]) {
  std::vector<Fr> st;
  for (uint32_t pc = vk->prog_off[prog]; pc < vk->prog_off[prog + 1]; pc++) {
    const uint32_t op = vk->ops[pc] & 0xff, arg = vk->ops[pc] >> 8;
    switch (op) {
      case OP_CONST: st.push_back(vk->constants[arg]); break;
      case OP_ADVICE: case OP_FIXED: case OP_INSTANCE: st.push_back(ev[op - OP_ADVICE][arg]); break;
      case OP_NEG: st.back() = fp_neg(st.back()); break;
      case OP_SCALE: st.back() = fp_mul(st.back(), vk->constants[arg]); break;
      default: {
        Fr b = st.back();
        st.pop_back();
        Fr& a = st.back();
        a = op == OP_ADD ? fp_add(a, b) : op == OP_SUB ? fp_sub(a, b) : fp_mul(a, b);
      }
    }
  }
  return st[0];
}

struct Domain {
  uint32_t k;
  uint64_t n;
  Fr omega, omega_inv;
  Fr rotate(const Fr& x, int32_t rot) const {
    return fp_mul(x, fr_pow_u64(rot >= 0 ? omega : omega_inv, (uint64_t)(rot >= 0 ? rot : -(int64_t)rot)));
  }
  // l_i(x) = w^i (x^n - 1) / (n (x - w^i)) for i in [lo, hi); negative i counts from the end of the domain
  std::vector<Fr> lagrange_range(const Fr& x, const Fr& xn, int64_t lo, int64_t hi) const {
    std::vector<Fr> out;
    const Fr num = fp_sub(xn, fr1()), nn = fp_from_u64<FrParams>(n);
    for (int64_t i = lo; i < hi; i++) {
      const Fr wi = rotate(fr1(), (int32_t)i);
      out.push_back(fp_mul(fp_mul(wi, num), fp_inv(fp_mul(nn, fp_sub(x, wi)))));
    }
    return out;
  }
};

}  // namespace

extern "C" {

int zg_vk_create(uint32_t k, const uint32_t* cs_words, size_t cs_nwords, const zg_fr* constants, size_t n_constants,
                 const zg_g1_affine* fixed_commitments, const zg_g1_affine* permutation_commitments,
                 const zg_fr* transcript_repr, zg_vk** out) {
  if (!out) return ZG_E_INVALID;
  *out = nullptr;
  if (!cs_words || !transcript_repr || k < 1 || k > 28 || (n_constants && !constants)) return ZG_E_INVALID;
  std::unique_ptr<zg_vk> vk(new zg_vk());
  vk->k = k;
  vk->constants.resize(n_constants);
  if (n_constants) memcpy(vk->constants.data(), constants, sizeof(Fr) * n_constants);
  int rc = vk_parse(vk.get(), cs_words, cs_nwords);
  if (rc) return rc;
  if ((vk->F && !fixed_commitments) || (vk->m && !permutation_commitments)) return ZG_E_INVALID;
  vk->fixed_comm.resize(vk->F);
  vk->perm_comm.resize(vk->m);
  if (vk->F) memcpy(vk->fixed_comm.data(), fixed_commitments, sizeof(G1Affine) * vk->F);
  if (vk->m) memcpy(vk->perm_comm.data(), permutation_commitments, sizeof(G1Affine) * vk->m);
  memcpy(vk->transcript_repr.v, transcript_repr, 32);
  *out = vk.release();
  return ZG_OK;
}

void zg_vk_free(zg_vk* vk) { delete vk; }
const char* zg_vk_last_error(const zg_vk* vk) { return vk ? vk->err.c_str() : "null vk"; }

int zg_pairing_check(const zg_g1_affine* p, const zg_g2_affine* q, size_t n, int* is_one) {
  if (!p || !q || !is_one) return ZG_E_INVALID;
  std::vector<G1Affine> P(n);
  std::vector<G2Aff> Q(n);
  memcpy(P.data(), p, sizeof(G1Affine) * n);
  for (size_t i = 0; i < n; i++) {
    memcpy(&Q[i].x, &q[i], sizeof(Fq2) * 2);
    Q[i].inf = f2_is_zero(Q[i].x) && f2_is_zero(Q[i].y);
    if (!g2_on_curve(Q[i]) || !(affine_is_identity(P[i]) || g1_on_curve(P[i]))) return ZG_E_INVALID;
  }
  *is_one = pairing_product_is_one(P, Q) ? 1 : 0;
  return ZG_OK;
}

int zg_verify_proof(zg_vk* vk, const zg_g1_affine* g1_generator, const zg_g2_affine* g2, const zg_g2_affine* s_g2,
                    const zg_fr* const* instances, const size_t* instance_lens, const uint8_t* proof, size_t proof_len) {
  if (!vk) return ZG_E_INVALID;
  if (!g1_generator || !g2 || !s_g2 || !proof || (vk->I && (!instances || !instance_lens)))
    return vk->fail(ZG_E_INVALID, "verify_proof: null argument");
  const uint32_t A = vk->A, I = vk->I, m = vk->m, Lk = vk->n_lookups, bf = vk->bf;
  const uint64_t n = 1ull << vk->k;
  const uint32_t chunk = vk->degree - 2, nsets = (m + chunk - 1) / chunk, qdeg = vk->degree - 1;
  Domain dom{vk->k, n, fr_root_of_unity(vk->k), fr0()};
  dom.omega_inv = fp_inv(dom.omega);
  std::vector<std::vector<Fr>> inst(I);
  for (uint32_t c = 0; c < I; c++) {
    if (instance_lens[c] > n - (bf + 1)) return vk->fail(ZG_E_INVALID, "verify_proof: instance too large");
    if (instance_lens[c] && !instances[c]) return vk->fail(ZG_E_INVALID, "verify_proof: null instance column");
    inst[c].resize(instance_lens[c]);
    if (instance_lens[c]) memcpy(inst[c].data(), instances[c], sizeof(Fr) * instance_lens[c]);
  }
  Reader tr{proof, proof_len};
  tr.common_scalar(vk->transcript_repr);
  for (uint32_t c = 0; c < I; c++)
    for (auto& v : inst[c]) tr.common_scalar(v);
  std::vector<G1Affine> adv_c(A), lk_a(Lk), lk_s(Lk), perm_z(nsets), lk_z(Lk), h_c(qdeg);
  for (auto& p : adv_c) p = tr.read_point();
  const Fr theta = tr.squeeze();
  for (uint32_t l = 0; l < Lk; l++) { lk_a[l] = tr.read_point(); lk_s[l] = tr.read_point(); }
  const Fr beta = tr.squeeze();
  const Fr gamma = tr.squeeze();
  for (auto& p : perm_z) p = tr.read_point();
  for (auto& p : lk_z) p = tr.read_point();
  const G1Affine random_c = tr.read_point();
  const Fr y = tr.squeeze();
  for (auto& p : h_c) p = tr.read_point();
  const Fr x = tr.squeeze();
  if (!tr.ok) return vk->fail(ZG_E_VERIFY, "verify_proof: malformed proof (commitments)");
  const Fr xn = fr_pow_u64(x, n);
  std::vector<Fr> ev[3];
  for (size_t i = 0; i < vk->q[0].size(); i++) ev[0].push_back(tr.read_scalar());
  // instance evaluations from the public inputs (KZG: QUERY_INSTANCE = false)
  for (auto& qi : vk->q[2]) {
    const std::vector<Fr>& vals = inst[qi.first];
    std::vector<Fr> ls = dom.lagrange_range(x, xn, -(int64_t)qi.second, (int64_t)vals.size() - qi.second);
    Fr acc = fr0();
    for (size_t i = 0; i < vals.size(); i++) acc = fp_add(acc, fp_mul(vals[i], ls[i]));
    ev[2].push_back(acc);
  }
  for (size_t i = 0; i < vk->q[1].size(); i++) ev[1].push_back(tr.read_scalar());
  const Fr random_eval = tr.read_scalar();
  std::vector<Fr> sigma_ev(m);
  for (auto& s : sigma_ev) s = tr.read_scalar();
  struct SetEv { Fr z, z_next, z_last; };
  std::vector<SetEv> sets(nsets);
  for (uint32_t i = 0; i < nsets; i++) {
    sets[i].z = tr.read_scalar();
    sets[i].z_next = tr.read_scalar();
    sets[i].z_last = (i + 1 < nsets) ? tr.read_scalar() : fr0();
  }
  struct LkEv { Fr z, z_next, a, a_inv, s; };
  std::vector<LkEv> lks(Lk);
  for (auto& l : lks) { l.z = tr.read_scalar(); l.z_next = tr.read_scalar(); l.a = tr.read_scalar(); l.a_inv = tr.read_scalar(); l.s = tr.read_scalar(); }
  if (!tr.ok) return vk->fail(ZG_E_VERIFY, "verify_proof: malformed proof (evaluations)");

  // ---- expected h(x): all constraint expressions at x, folded by y, over the vanishing polynomial ----------------------------------
  std::vector<Fr> lv = dom.lagrange_range(x, xn, -(int64_t)(bf + 1), 1);
  const Fr l_last = lv[0], l_0 = lv[bf + 1];
  Fr l_blind = fr0();
  for (uint32_t i = 1; i <= bf; i++) l_blind = fp_add(l_blind, lv[i]);
  const Fr one = fr1(), active = fp_sub(one, fp_add(l_last, l_blind));
  Fr hacc = fr0();
  auto fold = [&](const Fr& e) { hacc = fp_add(fp_mul(hacc, y), e); };
  for (uint32_t g = 0; g < vk->n_gate_progs; g++) fold(eval_at_point(vk, g, ev));
  if (nsets) {
    fold(fp_mul(l_0, fp_sub(one, sets[0].z)));
    const Fr zl = sets[nsets - 1].z;
    fold(fp_mul(l_last, fp_sub(fp_sqr(zl), zl)));
    for (uint32_t i = 1; i < nsets; i++) fold(fp_mul(l_0, fp_sub(sets[i].z, sets[i - 1].z_last)));
    const Fr delta = fr_delta();
    for (uint32_t i = 0; i < nsets; i++) {
      Fr left = sets[i].z_next, right = sets[i].z;
      Fr cur = fp_mul(fp_mul(beta, x), fr_pow_u64(delta, (uint64_t)i * chunk));
      for (uint32_t j = i * chunk; j < std::min(m, (i + 1) * chunk); j++) {
        const uint32_t kind = vk->perm[j].first, col = vk->perm[j].second;
        int qi = -1;
        for (size_t t = 0; t < vk->q[kind].size(); t++)
          if (vk->q[kind][t].first == col && vk->q[kind][t].second == 0) { qi = (int)t; break; }
        if (qi < 0) return vk->fail(ZG_E_INVALID, "verify_proof: permutation column is not queried at the current row");
        const Fr v = ev[kind][qi];
        left = fp_mul(left, fp_add(fp_add(v, fp_mul(beta, sigma_ev[j])), gamma));
        right = fp_mul(right, fp_add(fp_add(v, cur), gamma));
        cur = fp_mul(cur, delta);
      }
      fold(fp_mul(fp_sub(left, right), active));
    }
  }
  for (uint32_t l = 0; l < Lk; l++) {
    Fr ci = fr0(), ct = fr0();
    for (uint32_t g = vk->in_first[l]; g < vk->in_first[l] + vk->in_count[l]; g++) ci = fp_add(fp_mul(ci, theta), eval_at_point(vk, g, ev));
    for (uint32_t g = vk->tab_first[l]; g < vk->tab_first[l] + vk->tab_count[l]; g++) ct = fp_add(fp_mul(ct, theta), eval_at_point(vk, g, ev));
    const LkEv& e = lks[l];
    fold(fp_mul(l_0, fp_sub(one, e.z)));
    fold(fp_mul(l_last, fp_sub(fp_sqr(e.z), e.z)));
    const Fr lhs = fp_mul(fp_mul(e.z_next, fp_add(e.a, beta)), fp_add(e.s, gamma));
    const Fr rhs = fp_mul(fp_mul(e.z, fp_add(ci, beta)), fp_add(ct, gamma));
    fold(fp_mul(fp_sub(lhs, rhs), active));
    fold(fp_mul(l_0, fp_sub(e.a, e.s)));
    fold(fp_mul(fp_mul(fp_sub(e.a, e.s), fp_sub(e.a, e.a_inv)), active));
  }
  const Fr xn_m1 = fp_sub(xn, one);
  if (fp_is_zero(xn_m1)) return vk->fail(ZG_E_VERIFY, "verify_proof: challenge x lies in the domain");
  const Fr expected_h = fp_mul(hacc, fp_inv(xn_m1));

  // ---- queries in the prover's order, grouped by opening point (first appearance) ------------------------------------------------------
  // Every query contributes (commitment, scalar) terms to the right-hand MSM; the h commitment is itself sum_i xn^i H_i.
  struct Q { std::vector<std::pair<const G1Affine*, Fr>> com; Fr eval; int32_t rot; };
  std::vector<Q> queries;
  auto single = [&](const G1Affine* c, const Fr& e, int32_t rot) { queries.push_back(Q{{{c, one}}, e, rot}); };
  for (size_t i = 0; i < vk->q[0].size(); i++) single(&adv_c[vk->q[0][i].first], ev[0][i], vk->q[0][i].second);
  for (uint32_t i = 0; i < nsets; i++) { single(&perm_z[i], sets[i].z, 0); single(&perm_z[i], sets[i].z_next, 1); }
  for (int i = (int)nsets - 2; i >= 0; i--) single(&perm_z[i], sets[i].z_last, -(int32_t)(bf + 1));
  for (uint32_t l = 0; l < Lk; l++) {
    single(&lk_z[l], lks[l].z, 0);
    single(&lk_a[l], lks[l].a, 0);
    single(&lk_s[l], lks[l].s, 0);
    single(&lk_a[l], lks[l].a_inv, -1);
    single(&lk_z[l], lks[l].z_next, 1);
  }
  for (size_t i = 0; i < vk->q[1].size(); i++) single(&vk->fixed_comm[vk->q[1][i].first], ev[1][i], vk->q[1][i].second);
  for (uint32_t c = 0; c < m; c++) single(&vk->perm_comm[c], sigma_ev[c], 0);
  {
    Q hq;
    Fr pw = one;
    for (uint32_t i = 0; i < qdeg; i++) { hq.com.push_back({&h_c[i], pw}); pw = fp_mul(pw, xn); }
    hq.eval = expected_h;
    hq.rot = 0;
    queries.push_back(hq);
  }
  single(&random_c, random_eval, 0);
  std::vector<int32_t> rots;
  for (auto& qq : queries)
    if (std::find(rots.begin(), rots.end(), qq.rot) == rots.end()) rots.push_back(qq.rot);
  const Fr v = tr.squeeze();
  std::vector<G1Affine> ws(rots.size());
  for (auto& w : ws) w = tr.read_point();
  const Fr u = tr.squeeze();
  if (!tr.ok || tr.pos != proof_len) return vk->fail(ZG_E_VERIFY, "verify_proof: malformed proof (openings / trailing bytes)");

  // left = sum_i u^i W_i (paired with [s]G2);  right = sum_i u^i (sum_j v^j C_ij - e_i G + z_i W_i) (paired with -G2)
  std::vector<G1Affine> lp, rp;
  std::vector<Fr> ls, rs;
  Fr pu = one, eval_total = fr0();
  for (size_t g = 0; g < rots.size(); g++) {
    const Fr z = dom.rotate(x, rots[g]);
    Fr pw = one, eacc = fr0();
    for (auto& qq : queries) {
      if (qq.rot != rots[g]) continue;
      for (auto& term : qq.com) {
        rp.push_back(*term.first);
        rs.push_back(fp_mul(fp_mul(term.second, pw), pu));
      }
      eacc = fp_add(eacc, fp_mul(qq.eval, pw));
      pw = fp_mul(pw, v);
    }
    eval_total = fp_add(eval_total, fp_mul(eacc, pu));
    rp.push_back(ws[g]);
    rs.push_back(fp_mul(z, pu));
    lp.push_back(ws[g]);
    ls.push_back(pu);
    pu = fp_mul(pu, u);
  }
  G1Affine gen;
  memcpy(&gen, g1_generator, sizeof(gen));
  if (!g1_on_curve(gen)) return vk->fail(ZG_E_INVALID, "verify_proof: bad G1 generator");
  rp.push_back(gen);
  rs.push_back(fp_neg(eval_total));
  const G1Affine left = g1_to_affine(g1_msm(lp, ls));
  G1Affine right = g1_to_affine(g1_msm(rp, rs));
  if (!affine_is_identity(right)) right.y = fp_neg(right.y);     // e(left, [s]G2) * e(-right, G2) == 1
  std::vector<G2Aff> Qs(2);
  memcpy(&Qs[0].x, s_g2, sizeof(Fq2) * 2);
  memcpy(&Qs[1].x, g2, sizeof(Fq2) * 2);
  for (auto& qq : Qs) {
    qq.inf = f2_is_zero(qq.x) && f2_is_zero(qq.y);
    if (qq.inf || !g2_on_curve(qq)) return vk->fail(ZG_E_INVALID, "verify_proof: bad G2 point in the parameters");
  }
  if (!pairing_product_is_one({left, right}, Qs)) return vk->fail(ZG_E_VERIFY, "verify_proof: pairing check failed");
  return ZG_OK;
}

}  // extern "C"
