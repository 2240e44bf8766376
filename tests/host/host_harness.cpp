// CPU-side harness: compiles the device headers' portable paths with g++ so the shared
// field/curve arithmetic can be checked against Python big integers without a GPU.
// Test infrastructure only (never linked into libzg_b200.so).
#include "../../0g-halo2_b200/csrc/curve.cuh"
using namespace zg;
extern "C" {
void h_fr_mul(const Fr* a, const Fr* b, Fr* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_mul(a[i], b[i]); }
void h_fq_mul(const Fq* a, const Fq* b, Fq* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_mul(a[i], b[i]); }
void h_fr_add(const Fr* a, const Fr* b, Fr* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_add(a[i], b[i]); }
void h_fr_sub(const Fr* a, const Fr* b, Fr* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_sub(a[i], b[i]); }
void h_fq_add(const Fq* a, const Fq* b, Fq* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_add(a[i], b[i]); }
void h_fq_sub(const Fq* a, const Fq* b, Fq* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_sub(a[i], b[i]); }
void h_fr_inv(const Fr* a, Fr* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_inv(a[i]); }
void h_fq_inv(const Fq* a, Fq* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_inv(a[i]); }
// the 8 x 32-bit-limb body (device multiplier variant 0), which the host no longer uses for fp_mul
void h_fr_mul_portable(const Fr* a, const Fr* b, Fr* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_mul_portable(a[i], b[i]); }
void h_fq_mul_portable(const Fq* a, const Fq* b, Fq* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_mul_portable(a[i], b[i]); }
void h_fr_to_mont(const Fr* a, Fr* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_to_mont(a[i]); }
void h_fr_from_mont(const Fr* a, Fr* o, int n) { for (int i = 0; i < n; i++) o[i] = fp_from_mont(a[i]); }
// acc (xyzz) += affine points one by one; returns Jacobian
void h_madd_chain(const G1Affine* pts, int n, G1Jac* out) {
  G1Xyzz acc = xyzz_identity();
  for (int i = 0; i < n; i++) if (!affine_is_identity(pts[i])) xyzz_madd(acc, pts[i].x, pts[i].y);
  *out = xyzz_to_jacobian(acc);
}
// tree-ish: sum of pairwise (xyzz_add of two madd-chains)
void h_add_two_chains(const G1Affine* p, int n1, const G1Affine* q, int n2, G1Jac* out) {
  G1Xyzz a = xyzz_identity(), b = xyzz_identity();
  for (int i = 0; i < n1; i++) if (!affine_is_identity(p[i])) xyzz_madd(a, p[i].x, p[i].y);
  for (int i = 0; i < n2; i++) if (!affine_is_identity(q[i])) xyzz_madd(b, q[i].x, q[i].y);
  xyzz_add(a, b);
  *out = xyzz_to_jacobian(a);
}
}
