// EXPERIMENTAL (not used by any kernel yet): Montgomery multiplication on the FP64 pipe.
//
// B200's FP64 pipe issues 64 DFMA/clk/SM -- 2.2 to 2.7 times the rate of the IMAD.WIDE instructions that bound every
// kernel of this path today (profiles/r01_notes.md, "Pipe microbenchmarks").  Following Emmart, Zheng and Weems
// ("Faster modular exponentiation using double precision floating point arithmetic on the GPU", ARITH 2018), a field
// element is held as five 52-bit limbs in doubles and the Montgomery radix becomes 2^260:
//     mul(a, b) = a * b * 2^-260 mod p
// The exact 104-bit product of two limbs is obtained with two fused multiply-adds in round-toward-zero mode:
//     hi = fma_rz(a, b, 2^104)                 = 2^104 + floor(ab / 2^52) * 2^52      (the low 52 bits are truncated)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)   = 2^52 + (ab mod 2^52)                 (exact)
// so that the mantissa fields of `hi` and `lo` ARE the two 52-bit halves of the product.  This header is the
// reference formulation (clarity first) that tests/test_host_models.py checks against Python big integers on the CPU
// (IEEE-754 makes the host's fma under FE_TOWARDZERO bit-identical to the device's __fma_rz); the device port with raw
// bit-pattern accumulation, and the change of Montgomery radix across the library, are round-2 work.
#pragma once
#include <stdint.h>
#include <string.h>
#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

namespace zg52 {

constexpr uint64_t M52 = (1ull << 52) - 1;

struct Params52 {
  double p[5];        // modulus, 52-bit limbs
  uint64_t pinv;      // -p^-1 mod 2^52
};

struct Fe52 {
  double l[5];        // exact integers in [0, 2^52); value = sum l[i] * 2^(52 i)
};

#if defined(__CUDA_ARCH__)
#define ZG52_HD __host__ __device__ inline
__device__ __forceinline__ double fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
#else
#define ZG52_HD inline
// the caller has set FE_TOWARDZERO (tests/host/host_field52.cpp); std::fma is then the device's __fma_rz
inline double fma_rz(double a, double b, double c) { return std::fma(a, b, c); }
#endif

inline uint64_t bits_of(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}

// the two 52-bit halves of a * b (a, b exact integers below 2^52)
inline void two_prod(double a, double b, uint64_t& hi, uint64_t& lo) {
  const double C1 = 20282409603651670423947251286016.0;                 // 2^104
  const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;   // 2^104 + 2^52
  const double h = fma_rz(a, b, C1);
  const double l = fma_rz(a, b, C2 - h);
  hi = bits_of(h) & M52;
  lo = bits_of(l) & M52;
}

// a * b * 2^-260 mod p, fully reduced
inline Fe52 mul(const Fe52& a, const Fe52& b, const Params52& P) {
  uint64_t c[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 5; i++) {
    for (int j = 0; j < 5; j++) {
      uint64_t h, l;
      two_prod(a.l[j], b.l[i], h, l);
      c[i + j] += l;
      c[i + j + 1] += h;
    }
    // column i is complete: its low 52 bits decide the multiple of p that clears it
    const uint64_t low = c[i] & M52;
    c[i + 1] += c[i] >> 52;
    const uint64_t q = (low * P.pinv) & M52;
    const double qd = (double)q;                                         // exact: q < 2^52
    uint64_t h0, l0;
    two_prod(qd, P.p[0], h0, l0);
    c[i + 1] += h0 + ((low + l0) >> 52);                                 // low + l0 == 0 mod 2^52 by construction
    for (int j = 1; j < 5; j++) {
      uint64_t h, l;
      two_prod(qd, P.p[j], h, l);
      c[i + j] += l;
      c[i + j + 1] += h;
    }
  }
  // columns 5..9 hold the result (< 2p) before carry propagation
  uint64_t r[5], carry = 0;
  for (int k = 0; k < 5; k++) {
    const uint64_t v = c[5 + k] + carry;
    r[k] = v & M52;
    carry = v >> 52;
  }
  // one conditional subtraction of p
  uint64_t d[5];
  int64_t borrow = 0;
  for (int k = 0; k < 5; k++) {
    int64_t v = (int64_t)r[k] - (int64_t)(uint64_t)P.p[k] + borrow;
    borrow = v >> 63;                                                    // -1 when negative
    d[k] = (uint64_t)v & M52;
  }
  const bool ge = (borrow == 0) || carry;                                // r >= p
  Fe52 out;
  for (int k = 0; k < 5; k++) out.l[k] = (double)(ge ? d[k] : r[k]);
  return out;
}

// 256-bit integer (four little-endian u64 words, < 2^260) <-> five 52-bit limbs
inline Fe52 from_words(const uint64_t w[4]) {
  Fe52 r;
  r.l[0] = (double)(w[0] & M52);
  r.l[1] = (double)(((w[0] >> 52) | (w[1] << 12)) & M52);
  r.l[2] = (double)(((w[1] >> 40) | (w[2] << 24)) & M52);
  r.l[3] = (double)(((w[2] >> 28) | (w[3] << 36)) & M52);
  r.l[4] = (double)(w[3] >> 16);
  return r;
}
inline void to_words(const Fe52& a, uint64_t w[4]) {
  const uint64_t l0 = (uint64_t)a.l[0], l1 = (uint64_t)a.l[1], l2 = (uint64_t)a.l[2], l3 = (uint64_t)a.l[3], l4 = (uint64_t)a.l[4];
  w[0] = l0 | (l1 << 52);
  w[1] = (l1 >> 12) | (l2 << 40);
  w[2] = (l2 >> 24) | (l3 << 28);
  w[3] = (l3 >> 36) | (l4 << 16);
}

}  // namespace zg52
