// Cooperative EC arithmetic for the latency-bound MSM tail: ONE XYZZ addition computed by four warps of a CTA.
//
// A lone warp issues one 4-cycle IMAD.WIDE at a time on its SM sub-partition, so a Montgomery product costs >= 544
// cycles of latency and the 14 products of an XYZZ addition ~5 us -- and the tail of an MSM batch (msm_tail.cu) is a
// chain of ~100 dependent additions.  The 14 products fall into FOUR dependency levels:
//     level 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1            (P = U2 - U1, R = S2 - S1)
//     level 2: PP = P^2, RR = R^2, ZZb = ZZ1 ZZ2, ZZZb = ZZZ1 ZZZ2
//     level 3: PPP = P PP, Q = U1 PP, ZZ3 = ZZb PP                              (X3 = RR - PPP - 2Q)
//     level 4: T1 = R (Q - X3), T2 = S1 PPP, ZZZ3 = ZZZb PPP                    (Y3 = T1 - T2)
// so a CTA of four warps -- four sub-partitions, four multiply pipes -- holds both operands REPLICATED in every warp
// (lane l of each warp works on the same addition l), warp `role` computes product `role` of each level, and the
// products are exchanged through shared memory with one barrier per level: 4 product latencies instead of 14.  Doubling
// (9 products) takes 3 levels the same way.  Tree steps between lanes (shuffles) act on the replicated values and are
// simply executed by all four warps.
//
// The level functions below are plain ZG_HD code shared by the device kernels (msm_tail_coop.cu) and by the host test
// (tests/host/coop_add_test.cpp), which runs the four roles in lockstep and compares with curve.cuh::xyzz_add / xyzz_double.
#pragma once
#include "curve.cuh"

namespace zg {

constexpr int COOP_ROLES = 4;

// ---- addition ------------------------------------------------------------------------------------------------------
struct CoopAddState {
  G1Xyzz a, b;                                   // the operands (replicated)
  Fq u1, s1, p, r, pp, rr, zzb, zzzb, ppp, q, x3, y3, zz3, zzz3;
};
constexpr int COOP_ADD_LEVELS = 4;

// the product `role` contributes at `level` (1-based); roles without work at a level return zero
ZG_HD Fq coop_add_compute(int level, int role, const CoopAddState& s) {
  switch (level * 4 + role) {
    case 4 + 0: return fp_mul(s.a.x, s.b.zz);
    case 4 + 1: return fp_mul(s.b.x, s.a.zz);
    case 4 + 2: return fp_mul(s.a.y, s.b.zzz);
    case 4 + 3: return fp_mul(s.b.y, s.a.zzz);
    case 8 + 0: return fp_sqr(s.p);
    case 8 + 1: return fp_sqr(s.r);
    case 8 + 2: return fp_mul(s.a.zz, s.b.zz);
    case 8 + 3: return fp_mul(s.a.zzz, s.b.zzz);
    case 12 + 0: return fp_mul(s.p, s.pp);
    case 12 + 1: return fp_mul(s.u1, s.pp);
    case 12 + 2: return fp_mul(s.zzb, s.pp);
    case 16 + 0: return fp_mul(s.r, fp_sub(s.q, s.x3));
    case 16 + 1: return fp_mul(s.s1, s.ppp);
    case 16 + 2: return fp_mul(s.zzzb, s.ppp);
    default: return fp_zero<FqParams>();
  }
}
// all four roles' products of `level` -> the replicated state
ZG_HD void coop_add_absorb(int level, CoopAddState& s, const Fq (&o)[COOP_ROLES]) {
  switch (level) {
    case 1:
      s.u1 = o[0];
      s.s1 = o[2];
      s.p = fp_sub(o[1], o[0]);
      s.r = fp_sub(o[3], o[2]);
      break;
    case 2:
      s.pp = o[0];
      s.rr = o[1];
      s.zzb = o[2];
      s.zzzb = o[3];
      break;
    case 3:
      s.ppp = o[0];
      s.q = o[1];
      s.zz3 = o[2];
      s.x3 = fp_sub(fp_sub(s.rr, s.ppp), fp_dbl(s.q));
      break;
    default:
      s.zzz3 = o[2];
      s.y3 = fp_sub(o[0], o[1]);
      break;
  }
}
// does this (replicated) pair need the generic formulas at all?
ZG_HD bool coop_add_generic(const G1Xyzz& a, const G1Xyzz& b, bool active) {
  return active && !xyzz_is_identity(a) && !xyzz_is_identity(b);
}
// the sum after level 4 (or without any level when coop_add_generic is false), every special case of xyzz_add included
ZG_HD G1Xyzz coop_add_result(const CoopAddState& s, bool active) {
  if (!active || xyzz_is_identity(s.b)) return s.a;
  if (xyzz_is_identity(s.a)) return s.b;
  if (fp_is_zero(s.p)) return fp_is_zero(s.r) ? xyzz_double(s.a) : xyzz_identity();   // same x: doubling (rare) or inverse
  G1Xyzz o;
  o.x = s.x3;
  o.y = s.y3;
  o.zz = s.zz3;
  o.zzz = s.zzz3;
  return o;
}

// ---- doubling (dbl-2008-s-1, a = 0) ------------------------------------------------------------------------------------
struct CoopDblState {
  G1Xyzz a;
  Fq u, v, w, s, m, x3, y3, zz3, zzz3;
};
constexpr int COOP_DBL_LEVELS = 3;

ZG_HD Fq coop_dbl_compute(int level, int role, const CoopDblState& s) {
  switch (level * 4 + role) {
    case 4 + 0: return fp_sqr(fp_dbl(s.a.y));                 // V = (2Y)^2
    case 4 + 1: return fp_sqr(s.a.x);                         // XX
    case 8 + 0: return fp_mul(s.u, s.v);                      // W = U V
    case 8 + 1: return fp_mul(s.a.x, s.v);                    // S = X V
    case 8 + 2: return fp_mul(s.v, s.a.zz);                   // ZZ3
    case 8 + 3: return fp_sqr(s.m);                           // M^2
    case 12 + 0: return fp_mul(s.m, fp_sub(s.s, s.x3));
    case 12 + 1: return fp_mul(s.w, s.a.y);
    case 12 + 2: return fp_mul(s.w, s.a.zzz);                 // ZZZ3
    default: return fp_zero<FqParams>();
  }
}
ZG_HD void coop_dbl_absorb(int level, CoopDblState& s, const Fq (&o)[COOP_ROLES]) {
  switch (level) {
    case 1:
      s.u = fp_dbl(s.a.y);
      s.v = o[0];
      s.m = fp_add(fp_dbl(o[1]), o[1]);
      break;
    case 2:
      s.w = o[0];
      s.s = o[1];
      s.zz3 = o[2];
      s.x3 = fp_sub(o[3], fp_dbl(o[1]));
      break;
    default:
      s.zzz3 = o[2];
      s.y3 = fp_sub(o[0], o[1]);
      break;
  }
}
ZG_HD G1Xyzz coop_dbl_result(const CoopDblState& s) {
  if (xyzz_is_identity(s.a)) return s.a;
  G1Xyzz o;
  o.x = s.x3;
  o.y = s.y3;
  o.zz = s.zz3;
  o.zzz = s.zzz3;
  return o;
}

}  // namespace zg
