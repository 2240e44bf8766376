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

// acc += (x2, y2) affine, non-identity affine operand (madd-2008-s)
ZG_HD void xyzz_madd(G1Xyzz& acc, const Fq& x2, const Fq& y2) {
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
  Fq pp = fp_sqr(p);
  Fq ppp = fp_mul(p, pp);
  Fq q = fp_mul(acc.x, pp);
  Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
  Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(acc.y, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = fp_mul(acc.zz, pp);
  acc.zzz = fp_mul(acc.zzz, ppp);
}

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

#if defined(__CUDACC__)
// Latency-oriented forms of the two group operations for single-warp chains (msm_tail.cu): the same formulas as xyzz_add /
// xyzz_double with their independent products issued three at a time (fp_mul3_outlined) -- 5 calls instead of 14 products
// for an addition, 4 instead of 9 for a doubling.
__device__ __forceinline__ void xyzz_add_ilp(G1Xyzz& acc, const G1Xyzz& b) {
  if (xyzz_is_identity(b)) return;
  if (xyzz_is_identity(acc)) {
    acc = b;
    return;
  }
  Fq A[3], B[3], R[3];
  A[0] = acc.x; B[0] = b.zz;  A[1] = b.x; B[1] = acc.zz;  A[2] = acc.y; B[2] = b.zzz;
  fp_mul3_outlined<FqParams>(A, B, R);
  const Fq u1 = R[0], u2 = R[1], s1 = R[2];
  A[0] = b.y; B[0] = acc.zzz;  A[1] = acc.zz; B[1] = b.zz;  A[2] = acc.zzz; B[2] = b.zzz;
  fp_mul3_outlined<FqParams>(A, B, R);
  const Fq s2 = R[0], zz12 = R[1], zzz12 = R[2];
  const Fq p = fp_sub(u2, u1), r = fp_sub(s2, s1);
  if (fp_is_zero(p)) {
    if (fp_is_zero(r)) acc = xyzz_double(acc);
    else acc = xyzz_identity();
    return;
  }
  A[0] = p; B[0] = p;  A[1] = r; B[1] = r;  A[2] = p; B[2] = p;
  fp_mul3_outlined<FqParams>(A, B, R);
  const Fq pp = R[0], rr = R[1];
  A[0] = p; B[0] = pp;  A[1] = u1; B[1] = pp;  A[2] = zz12; B[2] = pp;
  fp_mul3_outlined<FqParams>(A, B, R);
  const Fq ppp = R[0], q = R[1];
  acc.zz = R[2];
  const Fq x3 = fp_sub(fp_sub(rr, ppp), fp_dbl(q));
  A[0] = r; B[0] = fp_sub(q, x3);  A[1] = s1; B[1] = ppp;  A[2] = zzz12; B[2] = ppp;
  fp_mul3_outlined<FqParams>(A, B, R);
  acc.x = x3;
  acc.y = fp_sub(R[0], R[1]);
  acc.zzz = R[2];
}
__device__ __forceinline__ G1Xyzz xyzz_double_ilp(const G1Xyzz& p) {
  if (xyzz_is_identity(p)) return p;
  G1Xyzz o;
  Fq A[3], B[3], R[3];
  const Fq u = fp_dbl(p.y);
  A[0] = u; B[0] = u;  A[1] = p.x; B[1] = p.x;  A[2] = u; B[2] = u;
  fp_mul3_outlined<FqParams>(A, B, R);
  const Fq v = R[0], xx = R[1];
  const Fq m = fp_add(fp_dbl(xx), xx);
  A[0] = u; B[0] = v;  A[1] = p.x; B[1] = v;  A[2] = v; B[2] = p.zz;
  fp_mul3_outlined<FqParams>(A, B, R);
  const Fq w = R[0], s = R[1];
  o.zz = R[2];
  A[0] = m; B[0] = m;  A[1] = w; B[1] = p.y;  A[2] = w; B[2] = p.zzz;
  fp_mul3_outlined<FqParams>(A, B, R);
  o.x = fp_sub(R[0], fp_dbl(s));
  const Fq wy = R[1];
  o.zzz = R[2];
  A[0] = m; B[0] = fp_sub(s, o.x);  A[1] = m; B[1] = m;  A[2] = m; B[2] = m;
  fp_mul3_outlined<FqParams>(A, B, R);
  o.y = fp_sub(R[0], wy);
  return o;
}
#endif

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
