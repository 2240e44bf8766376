// BN254 prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form R = 2^256.
//
// The byte layout of one element is identical to halo2curves' `Fr([u64; 4])` / `Fq([u64; 4])`
// (little-endian limbs, Montgomery form), so host buffers cross the C ABI unconverted.
// Replaces (on device) the field arithmetic of halo2curves tag 0.3.3 `bn256::{fr,fq}` that
// /root/reference/Cargo.toml:14-18 pins; call sites reach it through create_proof
// (/root/reference/src/wnn.rs:242-259).
//
// Multiplier bodies:
//   * a portable 64-bit-accumulator CIOS on 8 x 32-bit limbs (device variant 0; kept in step with the others by the tests),
//   * a row-wise PTX mad.lo.cc / madc.hi.cc CIOS (device only, ZG_MUL_VARIANT=1),
//   * the even/odd carry-chain CIOS (device only, ZG_MUL_VARIANT=2, the default): 120 IMAD.WIDE.U32(.X)
//     per product instead of 120 IMAD.WIDE + ~90 IMAD + ~240 IADD3 in what nvcc makes of the portable body,
//   * generated straight-line bodies for a dedicated squaring and for a*b + c*d under ONE Montgomery reduction
//     (field_gen.cuh, written and modelled instruction by instruction by scripts/gen_field_ops.py), and
//   * the HOST body: 4 x 64-bit limbs with 128-bit accumulators, plus 64-bit add / subtract and Kaliski's almost-inverse
//     (what the verifier, the witness synthesizer and the prover's per-round point normalisation run on).
// All compute the same functions bit for bit; tests/test_gpu_field.py (device) and tests/test_host_models.py (host) check
// them against Python big integers and each other.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZG_HD __host__ __device__ __forceinline__
#define ZG_D __device__ __forceinline__
#else
#define ZG_HD inline
#define ZG_D inline
#endif

// device multiplier: 0 = portable CIOS, 1 = row-wise mad.cc CIOS, 2 = even/odd carry chains (default)
#ifndef ZG_MUL_VARIANT
#define ZG_MUL_VARIANT 2
#endif

namespace zg {

struct FrParams {
  static constexpr uint32_t INV = 0xefffffffu;
  ZG_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  ZG_HD static constexpr uint32_t one(int i) {  // R mod r
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  ZG_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
  ZG_HD static constexpr uint32_t r3(int i) {  // R^3 mod r
    constexpr uint32_t m[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu,
                               0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
    return m[i];
  }
};

struct FqParams {
  static constexpr uint32_t INV = 0xe4866389u;
  ZG_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  ZG_HD static constexpr uint32_t one(int i) {
    constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  ZG_HD static constexpr uint32_t r2(int i) {
    constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
  ZG_HD static constexpr uint32_t r3(int i) {
    constexpr uint32_t m[8] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u,
                               0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u};
    return m[i];
  }
};

template <class P>
struct alignas(16) Fp {
  uint32_t v[8];
};
using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

template <class P>
ZG_HD Fp<P> fp_zero() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}
template <class P>
ZG_HD Fp<P> fp_one() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
  return r;
}
template <class P>
ZG_HD bool fp_is_zero(const Fp<P>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}
template <class P>
ZG_HD bool fp_eq(const Fp<P>& a, const Fp<P>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

#if !defined(__CUDA_ARCH__)
// Host-side view of an element as 4 x 64-bit limbs (same bytes on a little-endian host): the host arithmetic below
// (verifier, witness synthesis, per-round point normalisation) runs on it.
template <class P>
struct FpHost64 {
  static constexpr uint64_t mod(int i) { return (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
  // -p^-1 mod 2^64 from the 32-bit constant by one Newton step
  static constexpr uint64_t inv() {
    uint64_t x = (uint64_t)(0u - P::INV);          // p^-1 mod 2^32
    x = x * (2 - mod(0) * x);                      // p^-1 mod 2^64
    return 0 - x;
  }
};
template <class P>
inline void fp_load64(const Fp<P>& a, uint64_t (&x)[4]) {
  for (int i = 0; i < 4; i++) x[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
template <class P>
inline Fp<P> fp_store64(const uint64_t (&x)[4]) {
  Fp<P> r;
  for (int i = 0; i < 4; i++) {
    r.v[2 * i] = (uint32_t)x[i];
    r.v[2 * i + 1] = (uint32_t)(x[i] >> 32);
  }
  return r;
}
// x + y mod p and x - y mod p on the 64-bit view (x, y < p)
template <class P>
inline Fp<P> fp_add_host64(const Fp<P>& a, const Fp<P>& b) {
  typedef unsigned __int128 u128;
  uint64_t x[4], y[4], t[4], u[4];
  fp_load64(a, x);
  fp_load64(b, y);
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)x[i] + y[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)t[i] - FpHost64<P>::mod(i) - (uint64_t)bw;
    u[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  return bw ? fp_store64<P>(t) : fp_store64<P>(u);   // a + b < 2p < 2^255: no carry out of limb 3
}
template <class P>
inline Fp<P> fp_sub_host64(const Fp<P>& a, const Fp<P>& b) {
  typedef unsigned __int128 u128;
  uint64_t x[4], y[4], t[4];
  fp_load64(a, x);
  fp_load64(b, y);
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)x[i] - y[i] - (uint64_t)bw;
    t[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  const uint64_t mask = 0 - (uint64_t)bw;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)t[i] + (FpHost64<P>::mod(i) & mask);
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  return fp_store64<P>(t);
}
#endif

// r = (t >= p) ? t - p : t, for t < 2p
template <class P>
ZG_HD void fp_final_sub(uint32_t (&t)[8]) {
  uint32_t u[8];
#if defined(__CUDA_ARCH__)
  uint32_t borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]),
        "=r"(u[7]), "=r"(borrow)
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "n"(P::mod(0)), "n"(P::mod(1)), "n"(P::mod(2)), "n"(P::mod(3)), "n"(P::mod(4)),
        "n"(P::mod(5)), "n"(P::mod(6)), "n"(P::mod(7)));
  bool keep = borrow != 0;  // t < p
#else
  uint64_t bw = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)t[i] - P::mod(i) - bw;
    u[i] = (uint32_t)d;
    bw = (d >> 63) & 1;
  }
  bool keep = bw != 0;
#endif
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = keep ? t[i] : u[i];
}

template <class P>
ZG_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
#if !defined(__CUDA_ARCH__)
  return fp_add_host64<P>(a, b);
#else
  uint32_t t[8];
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]),
        "=r"(t[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
        "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    t[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  fp_final_sub<P>(t);  // a + b < 2p < 2^256: no carry out of limb 7
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  return r;
#endif
}

template <class P>
ZG_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
#if !defined(__CUDA_ARCH__)
  return fp_sub_host64<P>(a, b);
#else
  uint32_t t[8];
  Fp<P> r;
#if defined(__CUDA_ARCH__)
  uint32_t borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]),
        "=r"(t[7]), "=r"(borrow)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
        "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // add back p masked by the borrow (borrow is 0 or 0xffffffff)
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(P::mod(0) & borrow), "r"(P::mod(1) & borrow), "r"(P::mod(2) & borrow),
        "r"(P::mod(3) & borrow), "r"(P::mod(4) & borrow), "r"(P::mod(5) & borrow),
        "r"(P::mod(6) & borrow), "r"(P::mod(7) & borrow));
#else
  uint64_t bw = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)a.v[i] - b.v[i] - bw;
    t[i] = (uint32_t)d;
    bw = (d >> 63) & 1;
  }
  uint32_t mask = bw ? 0xffffffffu : 0u;
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)t[i] + (P::mod(i) & mask);
    r.v[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return r;
#endif
}

template <class P>
ZG_HD Fp<P> fp_neg(const Fp<P>& a) {
  return fp_sub<P>(fp_zero<P>(), a);
}
template <class P>
ZG_HD Fp<P> fp_dbl(const Fp<P>& a) {
  return fp_add<P>(a, a);
}

// ---- Montgomery multiplication -------------------------------------------------------
// Portable fused CIOS ("no-carry" form: valid because the top modulus limb is < 2^31 - 1, so
// the two per-row carries sum without overflow).  Inputs < p, output < p.
template <class P>
ZG_HD Fp<P> fp_mul_portable(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t A = (uint64_t)a.v[0] * b.v[i] + t[0];
    uint32_t m = (uint32_t)A * P::INV;
    uint64_t C = (uint64_t)m * P::mod(0) + (uint32_t)A;
    A >>= 32;
    C >>= 32;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      A += (uint64_t)a.v[j] * b.v[i] + t[j];
      C += (uint64_t)m * P::mod(j) + (uint32_t)A;
      A >>= 32;
      t[j - 1] = (uint32_t)C;
      C >>= 32;
    }
    t[7] = (uint32_t)(A + C);
  }
  fp_final_sub<P>(t);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  return r;
}

#if !defined(__CUDA_ARCH__)
// Host multiplier: the same function on 4 x 64-bit limbs with 128-bit accumulators (same bytes in memory on a
// little-endian host).  The host-side verifier (pairing, GWC multi-scalar product), the witness synthesizer and the
// per-round point normalisation of the prover run on it; about five times the speed of the 32-bit-limb body above.
template <class P>
inline Fp<P> fp_mul_host64(const Fp<P>& a, const Fp<P>& b) {
  typedef unsigned __int128 u128;
  uint64_t A[4], B[4], t[5] = {0, 0, 0, 0, 0};
  fp_load64(a, A);
  fp_load64(b, B);
  constexpr uint64_t M0 = FpHost64<P>::mod(0), M1 = FpHost64<P>::mod(1), M2 = FpHost64<P>::mod(2), M3 = FpHost64<P>::mod(3);
  constexpr uint64_t INV = FpHost64<P>::inv();
  for (int i = 0; i < 4; i++) {
    u128 c = (u128)A[0] * B[i] + t[0];
    t[0] = (uint64_t)c;
    c = (c >> 64) + (u128)A[1] * B[i] + t[1];
    t[1] = (uint64_t)c;
    c = (c >> 64) + (u128)A[2] * B[i] + t[2];
    t[2] = (uint64_t)c;
    c = (c >> 64) + (u128)A[3] * B[i] + t[3];
    t[3] = (uint64_t)c;
    c = (c >> 64) + t[4];
    t[4] = (uint64_t)c;                            // T < 2p * 2^64 < 2^319: no sixth limb
    const uint64_t m = t[0] * INV;
    c = (u128)m * M0 + t[0];
    c = (c >> 64) + (u128)m * M1 + t[1];
    t[0] = (uint64_t)c;
    c = (c >> 64) + (u128)m * M2 + t[2];
    t[1] = (uint64_t)c;
    c = (c >> 64) + (u128)m * M3 + t[3];
    t[2] = (uint64_t)c;
    c = (c >> 64) + t[4];
    t[3] = (uint64_t)c;
    t[4] = (uint64_t)(c >> 64);
  }
  // t < 2p < 2^255: t[4] == 0; one conditional subtraction
  uint64_t u[4];
  u128 bw = 0;
  const uint64_t M[4] = {M0, M1, M2, M3};
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)t[i] - M[i] - (uint64_t)bw;
    u[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  Fp<P> r;
  for (int i = 0; i < 4; i++) {
    const uint64_t x = bw ? t[i] : u[i];
    r.v[2 * i] = (uint32_t)x;
    r.v[2 * i + 1] = (uint32_t)(x >> 32);
  }
  return r;
}
#endif

#if defined(__CUDA_ARCH__)
// One CIOS row as a single asm statement, so the carry flag never crosses a statement
// boundary: t(9 limbs) += a * bi ; m = t0 * INV ; t += m * p ; (caller shifts by one limb).
template <class P>
__device__ __forceinline__ void mont_row_ptx(uint32_t (&t)[9], const Fp<P>& a, uint32_t bi) {
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      // low halves of a*bi
      "mad.lo.cc.u32 %0, %9, %17, %0;\n\t"
      "madc.lo.cc.u32 %1, %10, %17, %1;\n\t"
      "madc.lo.cc.u32 %2, %11, %17, %2;\n\t"
      "madc.lo.cc.u32 %3, %12, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %13, %17, %4;\n\t"
      "madc.lo.cc.u32 %5, %14, %17, %5;\n\t"
      "madc.lo.cc.u32 %6, %15, %17, %6;\n\t"
      "madc.lo.cc.u32 %7, %16, %17, %7;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      // high halves of a*bi
      "mad.hi.cc.u32 %1, %9, %17, %1;\n\t"
      "madc.hi.cc.u32 %2, %10, %17, %2;\n\t"
      "madc.hi.cc.u32 %3, %11, %17, %3;\n\t"
      "madc.hi.cc.u32 %4, %12, %17, %4;\n\t"
      "madc.hi.cc.u32 %5, %13, %17, %5;\n\t"
      "madc.hi.cc.u32 %6, %14, %17, %6;\n\t"
      "madc.hi.cc.u32 %7, %15, %17, %7;\n\t"
      "madc.hi.u32 %8, %16, %17, %8;\n\t"
      // m = t0 * (-p^-1 mod 2^32)
      "mul.lo.u32 m, %0, %18;\n\t"
      // low halves of m*p
      "mad.lo.cc.u32 %0, m, %19, %0;\n\t"
      "madc.lo.cc.u32 %1, m, %20, %1;\n\t"
      "madc.lo.cc.u32 %2, m, %21, %2;\n\t"
      "madc.lo.cc.u32 %3, m, %22, %3;\n\t"
      "madc.lo.cc.u32 %4, m, %23, %4;\n\t"
      "madc.lo.cc.u32 %5, m, %24, %5;\n\t"
      "madc.lo.cc.u32 %6, m, %25, %6;\n\t"
      "madc.lo.cc.u32 %7, m, %26, %7;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      // high halves of m*p
      "mad.hi.cc.u32 %1, m, %19, %1;\n\t"
      "madc.hi.cc.u32 %2, m, %20, %2;\n\t"
      "madc.hi.cc.u32 %3, m, %21, %3;\n\t"
      "madc.hi.cc.u32 %4, m, %22, %4;\n\t"
      "madc.hi.cc.u32 %5, m, %23, %5;\n\t"
      "madc.hi.cc.u32 %6, m, %24, %6;\n\t"
      "madc.hi.cc.u32 %7, m, %25, %7;\n\t"
      "madc.hi.u32 %8, m, %26, %8;\n\t"
      "}"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]),
        "+r"(t[7]), "+r"(t[8])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(bi), "n"(P::INV), "n"(P::mod(0)), "n"(P::mod(1)),
        "n"(P::mod(2)), "n"(P::mod(3)), "n"(P::mod(4)), "n"(P::mod(5)), "n"(P::mod(6)),
        "n"(P::mod(7)));
}

template <class P>
__device__ __forceinline__ Fp<P> fp_mul_ptx(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[9];
#pragma unroll
  for (int i = 0; i < 9; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    mont_row_ptx<P>(t, a, b.v[i]);
#pragma unroll
    for (int j = 0; j < 8; j++) t[j] = t[j + 1];  // t0 is zero after the row: shift one limb
    t[8] = 0;
  }
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = t[i];
  fp_final_sub<P>(u);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = u[i];
  return r;
}
#endif

#if defined(__CUDA_ARCH__)
// ---- even/odd carry-chain multiplier (the production device multiplier) -----------------------------
// The running CIOS accumulator T is kept as two 8-limb arrays, T = U + 2^32 * V.  Products of the EVEN
// limbs of a (or p) are 64-bit values aligned on U's limb pairs (0,1)(2,3)(4,5)(6,7); products of the ODD
// limbs are aligned on V's pairs.  Each group of four products is then ONE carry chain of
// mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into four IMAD.WIDE.U32(.X) with the carry in a
// predicate -- 16 wide multiply-adds per row and no separate carry-save adds.  After a row U[0] == 0;
// dividing by 2^32 renames the arrays: V becomes the new U, U >> 64 the new V, and the left-over limb
// U[1] enters the next row through one add whose carry feeds the first V chain.
// (validated limb for limb against Python big integers by tests/test_gpu_field.py)
template <class P>
__device__ __forceinline__ void eo_reduce(uint32_t (&u)[8], uint32_t (&w)[8]) {
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      "mul.lo.u32 m, %0, %16;\n\t"
      "mad.lo.cc.u32 %8, m, %18, %8;\n\t"
      "madc.hi.cc.u32 %9, m, %18, %9;\n\t"
      "madc.lo.cc.u32 %10, m, %20, %10;\n\t"
      "madc.hi.cc.u32 %11, m, %20, %11;\n\t"
      "madc.lo.cc.u32 %12, m, %22, %12;\n\t"
      "madc.hi.cc.u32 %13, m, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, m, %24, %14;\n\t"
      "madc.hi.u32 %15, m, %24, %15;\n\t"
      "mad.lo.cc.u32 %0, m, %17, %0;\n\t"
      "madc.hi.cc.u32 %1, m, %17, %1;\n\t"
      "madc.lo.cc.u32 %2, m, %19, %2;\n\t"
      "madc.hi.cc.u32 %3, m, %19, %3;\n\t"
      "madc.lo.cc.u32 %4, m, %21, %4;\n\t"
      "madc.hi.cc.u32 %5, m, %21, %5;\n\t"
      "madc.lo.cc.u32 %6, m, %23, %6;\n\t"
      "madc.hi.cc.u32 %7, m, %23, %7;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      "}"
      : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
        "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7])
      : "n"(P::INV), "n"(P::mod(0)), "n"(P::mod(1)), "n"(P::mod(2)), "n"(P::mod(3)), "n"(P::mod(4)),
        "n"(P::mod(5)), "n"(P::mod(6)), "n"(P::mod(7)));
}

template <class P>
__device__ __forceinline__ Fp<P> fp_mul_eo(const Fp<P>& a, const Fp<P>& b) {
  uint32_t u[8], w[8];
  // row 0: plain products
  asm("mul.lo.u32 %0, %16, %24;\n\t"
      "mul.hi.u32 %1, %16, %24;\n\t"
      "mul.lo.u32 %2, %18, %24;\n\t"
      "mul.hi.u32 %3, %18, %24;\n\t"
      "mul.lo.u32 %4, %20, %24;\n\t"
      "mul.hi.u32 %5, %20, %24;\n\t"
      "mul.lo.u32 %6, %22, %24;\n\t"
      "mul.hi.u32 %7, %22, %24;\n\t"
      "mul.lo.u32 %8, %17, %24;\n\t"
      "mul.hi.u32 %9, %17, %24;\n\t"
      "mul.lo.u32 %10, %19, %24;\n\t"
      "mul.hi.u32 %11, %19, %24;\n\t"
      "mul.lo.u32 %12, %21, %24;\n\t"
      "mul.hi.u32 %13, %21, %24;\n\t"
      "mul.lo.u32 %14, %23, %24;\n\t"
      "mul.hi.u32 %15, %23, %24;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(b.v[0]));
  eo_reduce<P>(u, w);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    // divide by 2^32: new U = V, new V = U >> 64, carry limb U[1]
    const uint32_t x1 = u[1];
    uint32_t nw[8];
#pragma unroll
    for (int j = 0; j < 6; j++) nw[j] = u[j + 2];
    nw[6] = 0;
    nw[7] = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      u[j] = w[j];
      w[j] = nw[j];
    }
    asm("add.cc.u32 %0, %0, %25;\n\t"
        "madc.lo.cc.u32 %8, %17, %24, %8;\n\t"
        "madc.hi.cc.u32 %9, %17, %24, %9;\n\t"
        "madc.lo.cc.u32 %10, %19, %24, %10;\n\t"
        "madc.hi.cc.u32 %11, %19, %24, %11;\n\t"
        "madc.lo.cc.u32 %12, %21, %24, %12;\n\t"
        "madc.hi.cc.u32 %13, %21, %24, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %24, %14;\n\t"
        "madc.hi.u32 %15, %23, %24, %15;\n\t"
        "mad.lo.cc.u32 %0, %16, %24, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
        "madc.lo.cc.u32 %2, %18, %24, %2;\n\t"
        "madc.hi.cc.u32 %3, %18, %24, %3;\n\t"
        "madc.lo.cc.u32 %4, %20, %24, %4;\n\t"
        "madc.hi.cc.u32 %5, %20, %24, %5;\n\t"
        "madc.lo.cc.u32 %6, %22, %24, %6;\n\t"
        "madc.hi.cc.u32 %7, %22, %24, %7;\n\t"
        "addc.u32 %15, %15, 0;"
        : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
          "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "r"(b.v[i]), "r"(x1));
    eo_reduce<P>(u, w);
  }
  // T = (U >> 32) + V  (< 2p), then one conditional subtraction
  uint32_t t[8];
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
      : "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
        "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]));
  fp_final_sub<P>(t);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  return r;
}
#endif

#if defined(__CUDA_ARCH__) && defined(ZG_FP_MUL_NOINLINE)
// Latency-bound kernels (one warp walking a chain of EC additions) want a SMALL instruction
// footprint: with the multiplier inlined a single xyzz_add is ~96 KB of straight-line SASS, far beyond
// the 32 KB L1.5 instruction cache, and a lone warp then runs at instruction-fetch speed.  Translation
// units that define ZG_FP_MUL_NOINLINE call one shared copy of the multiplier instead.
template <class P>
__device__ __noinline__ Fp<P> fp_mul_outlined(Fp<P> a, Fp<P> b) {
#if ZG_MUL_VARIANT == 0
  return fp_mul_portable<P>(a, b);
#else
  return fp_mul_eo<P>(a, b);
#endif
}
#endif

template <class P>
ZG_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && defined(ZG_FP_MUL_NOINLINE)
  return fp_mul_outlined<P>(a, b);
#elif defined(__CUDA_ARCH__) && ZG_MUL_VARIANT == 2
  return fp_mul_eo<P>(a, b);
#elif defined(__CUDA_ARCH__) && ZG_MUL_VARIANT == 1
  return fp_mul_ptx<P>(a, b);
#elif defined(__CUDA_ARCH__)
  return fp_mul_portable<P>(a, b);
#else
  return fp_mul_host64<P>(a, b);
#endif
}
template <class P>
ZG_HD Fp<P> fp_sqr(const Fp<P>& a) {
  return fp_mul<P>(a, a);
}

}  // namespace zg
// dedicated squaring and the two-product ("lazy reduction") multiply-add: generated straight-line PTX bodies, modelled
// instruction for instruction on the CPU (scripts/gen_field_ops.py, tests/test_field_gen.py)
#ifndef ZG_FIELD_GEN
#define ZG_FIELD_GEN 1
#endif
#include "field_gen.cuh"
namespace zg {

#if defined(__CUDA_ARCH__) && defined(ZG_FP_MUL_NOINLINE)
template <class P>
__device__ __noinline__ Fp<P> fp_sqr_gen_outlined(Fp<P> a) {
  return fp_sqr_gen<P>(a);
}
template <class P>
__device__ __noinline__ Fp<P> fp_mul2_gen_outlined(Fp<P> a, Fp<P> b, Fp<P> c, Fp<P> d) {
  return fp_mul2_gen<P>(a, b, c, d);
}
#endif

// a^2 with the dedicated squaring body on the device (100 wide products instead of 128)
template <class P>
ZG_HD Fp<P> fp_sqr_fast(const Fp<P>& a) {
#if defined(__CUDA_ARCH__) && ZG_FIELD_GEN && defined(ZG_FP_MUL_NOINLINE)
  return fp_sqr_gen_outlined<P>(a);
#elif defined(__CUDA_ARCH__) && ZG_FIELD_GEN
  return fp_sqr_gen<P>(a);
#else
  return fp_mul<P>(a, a);
#endif
}
// a*b + c*d with ONE Montgomery reduction on the device (192 wide products instead of 256); a, b, c, d <= p
template <class P>
ZG_HD Fp<P> fp_mul2add(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
#if defined(__CUDA_ARCH__) && ZG_FIELD_GEN && defined(ZG_FP_MUL_NOINLINE)
  return fp_mul2_gen_outlined<P>(a, b, c, d);
#elif defined(__CUDA_ARCH__) && ZG_FIELD_GEN
  return fp_mul2_gen<P>(a, b, c, d);
#else
  return fp_add<P>(fp_mul<P>(a, b), fp_mul<P>(c, d));
#endif
}
// -a as an OPERAND of fp_mul2add: p - a without the zero check (0 -> p, which the two-product body accepts)
template <class P>
ZG_HD Fp<P> fp_neg_lazy(const Fp<P>& a) {
#if defined(__CUDA_ARCH__) && ZG_FIELD_GEN
  Fp<P> r;
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "n"(P::mod(0)), "n"(P::mod(1)), "n"(P::mod(2)), "n"(P::mod(3)), "n"(P::mod(4)), "n"(P::mod(5)), "n"(P::mod(6)),
        "n"(P::mod(7)), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]));
  return r;
#else
  return fp_neg<P>(a);
#endif
}

// canonical (non-Montgomery) integer -> Montgomery form
template <class P>
ZG_HD Fp<P> fp_to_mont(const Fp<P>& raw) {
  Fp<P> r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.v[i] = P::r2(i);
  return fp_mul<P>(raw, r2);
}
// Montgomery form -> canonical integer limbs
template <class P>
ZG_HD Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> o = fp_zero<P>();
  o.v[0] = 1;
  return fp_mul<P>(a, o);
}
template <class P>
ZG_HD Fp<P> fp_from_u64(uint64_t x) {
  Fp<P> raw = fp_zero<P>();
  raw.v[0] = (uint32_t)x;
  raw.v[1] = (uint32_t)(x >> 32);
  return fp_to_mont<P>(raw);
}

// a^e for a 256-bit exponent given as 8 little-endian 32-bit limbs (square-and-multiply, MSB first)
template <class P>
ZG_HD Fp<P> fp_pow(const Fp<P>& a, const uint32_t* e, int nlimbs) {
  Fp<P> acc = fp_one<P>();
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      acc = fp_sqr<P>(acc);
      if ((e[i] >> b) & 1) acc = fp_mul<P>(acc, a);
    }
  }
  return acc;
}
template <class P>
ZG_HD Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t e) {
  uint32_t limbs[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
  return fp_pow<P>(a, limbs, 2);
}
// a^e, square-and-multiply from the highest set bit of e (variable time in e)
template <class P>
ZG_HD Fp<P> fp_pow_var(const Fp<P>& a, uint64_t e) {
  if (e == 0) return fp_one<P>();
  int top = 63;
  while (!((e >> top) & 1)) top--;
  Fp<P> acc = a;
  for (int b = top - 1; b >= 0; b--) {
    acc = fp_sqr<P>(acc);
    if ((e >> b) & 1) acc = fp_mul<P>(acc, a);
  }
  return acc;
}
// Fermat inverse a^(p-2); maps 0 -> 0 (same convention as ff::Field::invert().unwrap_or(0)
// inside halo2's batch_invert, which skips zeros).
#if !defined(__CUDA_ARCH__)
// Host inversion: Kaliski's almost-inverse (shifts, additions and subtractions on 256-bit integers, no modular step
// inside the loop) gives x = (a R)^-1 * 2^k mod p with 254 <= k <= 508; two Montgomery products by powers of two then
// land on a^-1 R.  About 3 us against 20 us for the 380 products of the Fermat exponentiation.  Same function as the
// device body below (the inverse is unique; 0 -> 0).
template <class P>
inline Fp<P> fp_inv_host(const Fp<P>& a) {
  typedef unsigned __int128 u128;
  if (fp_is_zero(a)) return a;
  const uint64_t M[4] = {FpHost64<P>::mod(0), FpHost64<P>::mod(1), FpHost64<P>::mod(2), FpHost64<P>::mod(3)};
  uint64_t u[4], v[4], r[4] = {0, 0, 0, 0}, s[4] = {1, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u[i] = M[i];
    v[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
  }
  auto shr1 = [](uint64_t* x) {
    x[0] = (x[0] >> 1) | (x[1] << 63);
    x[1] = (x[1] >> 1) | (x[2] << 63);
    x[2] = (x[2] >> 1) | (x[3] << 63);
    x[3] >>= 1;
  };
  auto shl1 = [](uint64_t* x) {
    x[3] = (x[3] << 1) | (x[2] >> 63);
    x[2] = (x[2] << 1) | (x[1] >> 63);
    x[1] = (x[1] << 1) | (x[0] >> 63);
    x[0] <<= 1;
  };
  auto gt = [](const uint64_t* x, const uint64_t* y) {
    for (int i = 3; i >= 0; i--)
      if (x[i] != y[i]) return x[i] > y[i];
    return false;
  };
  auto add = [](uint64_t* x, const uint64_t* y) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)x[i] + y[i];
      x[i] = (uint64_t)c;
      c >>= 64;
    }
  };
  auto sub = [](uint64_t* x, const uint64_t* y) {
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)x[i] - y[i] - (uint64_t)bw;
      x[i] = (uint64_t)d;
      bw = (d >> 64) & 1;
    }
  };
  uint32_t k = 0;
  while ((v[0] | v[1] | v[2] | v[3]) != 0) {      // r, s < 2p < 2^255 throughout
    if (!(u[0] & 1)) {
      shr1(u);
      shl1(s);
    } else if (!(v[0] & 1)) {
      shr1(v);
      shl1(r);
    } else if (gt(u, v)) {
      sub(u, v);
      shr1(u);
      add(r, s);
      shl1(s);
    } else {
      sub(v, u);
      shr1(v);
      add(s, r);
      shl1(r);
    }
    k++;
  }
  if (!gt(M, r)) sub(r, M);                         // r >= p
  uint64_t x[4] = {M[0], M[1], M[2], M[3]};
  sub(x, r);                                         // x = p - r = (aR)^-1 * 2^k mod p
  Fp<P> t, r2, pw;
  for (int i = 0; i < 4; i++) {
    t.v[2 * i] = (uint32_t)x[i];
    t.v[2 * i + 1] = (uint32_t)(x[i] >> 32);
  }
  for (int i = 0; i < 8; i++) r2.v[i] = P::r2(i);
  // a^-1 R = x * 2^(512 - k); mont(x, mont(R^2, 2^e)) = x * 2^e for 2^e < p (e <= 253)
  uint32_t e = 512 - k;
  auto pow2 = [&](uint32_t b) {
    Fp<P> o = fp_zero<P>();
    o.v[b >> 5] = 1u << (b & 31);
    return fp_mul_host64<P>(r2, o);                  // Montgomery form of 2^b
  };
  if (e > 253) {
    t = fp_mul_host64<P>(t, pow2(253));
    e -= 253;
  }
  return fp_mul_host64<P>(t, pow2(e));
}
#endif

template <class P>
ZG_HD Fp<P> fp_inv(const Fp<P>& a) {
#if !defined(__CUDA_ARCH__)
  return fp_inv_host<P>(a);
#else
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = P::mod(i);
  e[0] -= 2;  // low limb of both moduli is >= 2, no borrow
  return fp_pow<P>(a, e, 8);
#endif
}

}  // namespace zg
