// BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for the MSM kernels.
//
// Bucket accumulators use extended Jacobian ("XYZZ": x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2)
// coordinates: mixed add = 8M + 2S, full add = 12M + 2S.  Identity is ZZ = 0.
// Bases are affine {x, y} in the halo2curves `G1Affine` in-memory layout (identity = (0,0)),
// results leave the device as Jacobian {x, y, z} = halo2curves `G1` layout
// (replaces the group law behind `best_multiexp`, reached from
// /root/reference/src/wnn.rs:242-259 via ParamsKZG::commit / commit_lagrange).
#pragma once
#include "field.cuh"

namespace zg {

struct alignas(16) G1Affine {
  Fq x, y;
};
struct alignas(16) G1Jac {
  Fq x, y, z;
};
struct alignas(16) G1Xyzz {
  Fq x, y, zz, zzz;
};

ZG_HD bool affine_is_identity(const G1Affine& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
ZG_HD bool xyzz_is_identity(const G1Xyzz& p) { return fp_is_zero(p.zz); }

ZG_HD G1Xyzz xyzz_identity() {
  G1Xyzz r;
  r.x = fp_zero<FqParams>();
  r.y = fp_zero<FqParams>();
  r.zz = fp_zero<FqParams>();
  r.zzz = fp_zero<FqParams>();
  return r;
}

ZG_HD G1Xyzz xyzz_from_affine(const G1Affine& p) {
  G1Xyzz r;
  if (affine_is_identity(p)) return xyzz_identity();
  r.x = p.x;
  r.y = p.y;
  r.zz = fp_one<FqParams>();
  r.zzz = fp_one<FqParams>();
  return r;
}

// 2 * (affine point)  (dbl-2008-s-1 with ZZ1 = ZZZ1 = 1, a = 0)
ZG_HD G1Xyzz xyzz_double_affine(const G1Affine& p) {
  G1Xyzz r;
  Fq u = fp_dbl(p.y);
  Fq v = fp_sqr(u);
  Fq w = fp_mul(u, v);
  Fq s = fp_mul(p.x, v);
  Fq xx = fp_sqr(p.x);
  Fq m = fp_add(fp_dbl(xx), xx);
  r.x = fp_sub(fp_sqr(m), fp_dbl(s));
  r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * P  (dbl-2008-s-1, a = 0); P must not be the identity unless handled by the caller
ZG_HD G1Xyzz xyzz_double(const G1Xyzz& p) {
  if (xyzz_is_identity(p)) return p;
  G1Xyzz r;
  Fq u = fp_dbl(p.y);
  Fq v = fp_sqr(u);
  Fq w = fp_mul(u, v);
  Fq s = fp_mul(p.x, v);
  Fq xx = fp_sqr(p.x);
  Fq m = fp_add(fp_dbl(xx), xx);
  r.x = fp_sub(fp_sqr(m), fp_dbl(s));
  r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
  r.zz = fp_mul(v, p.zz);
  r.zzz = fp_mul(w, p.zzz);
  return r;
}

// acc += (x2, y2) affine, non-identity affine operand (madd-2008-s).
// LAZY: the two squarings use the dedicated squaring body and Y3 = R*(Q - X3) - Y1*PPP is ONE lazily reduced two-product
// (field.cuh: fp_sqr_fast / fp_mul2add): 9 Montgomery reductions and 1 096 wide products per addition instead of 10 and
// 1 280.  Both forms return the same canonical limbs.
template <bool LAZY>
ZG_HD void xyzz_madd_t(G1Xyzz& acc, const Fq& x2, const Fq& y2) {
  if (xyzz_is_identity(acc)) {
    acc.x = x2;
    acc.y = y2;
    acc.zz = fp_one<FqParams>();
    acc.zzz = fp_one<FqParams>();
    return;
  }
  Fq u2 = fp_mul(x2, acc.zz);
  Fq s2 = fp_mul(y2, acc.zzz);
  Fq p = fp_sub(u2, acc.x);
  Fq r = fp_sub(s2, acc.y);
  if (fp_is_zero(p)) {
    if (fp_is_zero(r)) {
      G1Affine a;
      a.x = x2;
      a.y = y2;
      acc = xyzz_double_affine(a);
    } else {
      acc = xyzz_identity();
    }
    return;
  }
  Fq pp = LAZY ? fp_sqr_fast(p) : fp_sqr(p);
  Fq ppp = fp_mul(p, pp);
  Fq q = fp_mul(acc.x, pp);
  Fq x3 = fp_sub(fp_sub(LAZY ? fp_sqr_fast(r) : fp_sqr(r), ppp), fp_dbl(q));
  Fq y3 = LAZY ? fp_mul2add(r, fp_sub(q, x3), fp_neg_lazy(acc.y), ppp)
               : fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(acc.y, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = fp_mul(acc.zz, pp);
  acc.zzz = fp_mul(acc.zzz, ppp);
}
ZG_HD void xyzz_madd(G1Xyzz& acc, const Fq& x2, const Fq& y2) { xyzz_madd_t<false>(acc, x2, y2); }

// acc += b (add-2008-s), all special cases handled
ZG_HD void xyzz_add(G1Xyzz& acc, const G1Xyzz& b) {
  if (xyzz_is_identity(b)) return;
  if (xyzz_is_identity(acc)) {
    acc = b;
    return;
  }
  Fq u1 = fp_mul(acc.x, b.zz);
  Fq u2 = fp_mul(b.x, acc.zz);
  Fq s1 = fp_mul(acc.y, b.zzz);
  Fq s2 = fp_mul(b.y, acc.zzz);
  Fq p = fp_sub(u2, u1);
  Fq r = fp_sub(s2, s1);
  if (fp_is_zero(p)) {
    if (fp_is_zero(r)) {
      acc = xyzz_double(acc);
    } else {
      acc = xyzz_identity();
    }
    return;
  }
  Fq pp = fp_sqr(p);
  Fq ppp = fp_mul(p, pp);
  Fq q = fp_mul(u1, pp);
  Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
  Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(s1, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = fp_mul(fp_mul(acc.zz, b.zz), pp);
  acc.zzz = fp_mul(fp_mul(acc.zzz, b.zzz), ppp);
}

// XYZZ -> Jacobian without an inversion: Z = ZZ*ZZZ, X' = X*ZZ*ZZZ^2, Y' = Y*ZZ^3*ZZZ^2
// (then X'/Z^2 = X/ZZ and Y'/Z^3 = Y/ZZZ).
ZG_HD G1Jac xyzz_to_jacobian(const G1Xyzz& p) {
  G1Jac r;
  if (xyzz_is_identity(p)) {
    r.x = fp_zero<FqParams>();
    r.y = fp_one<FqParams>();
    r.z = fp_zero<FqParams>();
    return r;
  }
  Fq z = fp_mul(p.zz, p.zzz);
  Fq t = fp_mul(z, p.zzz);           // ZZ * ZZZ^2
  r.x = fp_mul(p.x, t);
  r.y = fp_mul(fp_mul(p.y, t), fp_sqr(p.zz));
  r.z = z;
  return r;
}

}  // namespace zg
