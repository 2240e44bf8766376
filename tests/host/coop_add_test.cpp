// Host check of the cooperative (four-role) XYZZ addition and doubling of csrc/msm_coop.cuh: the four roles are run in
// lockstep, level by level, exactly as the four warps of a CTA run them between barriers (msm_tail_coop.cu), and the
// result every role ends with is compared with curve.cuh::xyzz_add / xyzz_double.  Built and run by tests/test_msm_coop_host.py.
#include <cstdio>
#include <random>
#include "msm_coop.cuh"
using namespace zg;

static std::mt19937_64 g(2024);
static Fq rnd() {
  Fq x;
  for (int i = 0; i < 8; i++) x.v[i] = (uint32_t)g();
  x.v[7] &= 0x1fffffffu;   // < 2^253 < q
  return x;
}
static G1Xyzz rnd_pt() { return {rnd(), rnd(), rnd(), rnd()}; }
static bool eq(const G1Xyzz& a, const G1Xyzz& b) {
  if (xyzz_is_identity(a) || xyzz_is_identity(b)) return xyzz_is_identity(a) && xyzz_is_identity(b);
  return fp_eq(a.x, b.x) && fp_eq(a.y, b.y) && fp_eq(a.zz, b.zz) && fp_eq(a.zzz, b.zzz);
}

static int check_add(const G1Xyzz& a, const G1Xyzz& b, bool active) {
  G1Xyzz expect = a;
  if (active) xyzz_add(expect, b);
  CoopAddState st[COOP_ROLES];
  for (int r = 0; r < COOP_ROLES; r++) { st[r].a = a; st[r].b = b; }
  if (coop_add_generic(a, b, active)) {
    for (int level = 1; level <= COOP_ADD_LEVELS; level++) {
      Fq out[COOP_ROLES];
      for (int r = 0; r < COOP_ROLES; r++) out[r] = coop_add_compute(level, r, st[r]);   // before the barrier
      for (int r = 0; r < COOP_ROLES; r++) coop_add_absorb(level, st[r], out);           // after it
    }
  }
  int bad = 0;
  for (int r = 0; r < COOP_ROLES; r++) bad += !eq(coop_add_result(st[r], active), expect);
  return bad;
}
static int check_dbl(const G1Xyzz& a) {
  G1Xyzz expect = xyzz_double(a);
  CoopDblState st[COOP_ROLES];
  for (int r = 0; r < COOP_ROLES; r++) st[r].a = a;
  for (int level = 1; level <= COOP_DBL_LEVELS; level++) {
    Fq out[COOP_ROLES];
    for (int r = 0; r < COOP_ROLES; r++) out[r] = coop_dbl_compute(level, r, st[r]);
    for (int r = 0; r < COOP_ROLES; r++) coop_dbl_absorb(level, st[r], out);
  }
  int bad = 0;
  for (int r = 0; r < COOP_ROLES; r++) bad += !eq(coop_dbl_result(st[r]), expect);
  return bad;
}

int main() {
  int bad = 0, n = 0;
  const G1Xyzz id = xyzz_identity();
  for (int it = 0; it < 2000; it++) {
    G1Xyzz a = rnd_pt(), b = rnd_pt();
    bad += check_add(a, b, true); n++;
    bad += check_add(a, b, false); n++;
    bad += check_dbl(a); n++;
  }
  for (int it = 0; it < 50; it++) {
    G1Xyzz a = rnd_pt();
    G1Xyzz neg = a;
    neg.y = fp_neg(a.y);
    // the same point in another representation: (X l^2, Y l^3, ZZ l^2, ZZZ l^3)
    Fq l = rnd(), l2 = fp_sqr(l), l3 = fp_mul(l2, l);
    G1Xyzz same = {fp_mul(a.x, l2), fp_mul(a.y, l3), fp_mul(a.zz, l2), fp_mul(a.zzz, l3)};
    G1Xyzz same_neg = same;
    same_neg.y = fp_neg(same.y);
    bad += check_add(a, id, true); bad += check_add(id, a, true); bad += check_add(id, id, true);
    bad += check_add(a, a, true); bad += check_add(a, neg, true);
    bad += check_add(a, same, true); bad += check_add(a, same_neg, true);
    bad += check_add(id, a, false); bad += check_add(a, a, false);
    bad += check_dbl(id);
    n += 10;
  }
  printf("coop add/dbl: %d checks, %d mismatches\n", n, bad);
  return bad != 0;
}
