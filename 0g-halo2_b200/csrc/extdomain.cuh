// The prover's INTERNAL extended evaluation domain.
//
// halo2_proofs (v2023_04_20, `poly/domain.rs`; reached from /root/reference/src/wnn.rs:242-259) evaluates the quotient
// numerator on the coset zeta * <omega_ext> of size 2^extended_k >= n * (degree - 1): 8n for zero_g's degree-6 system.
// None of that leaves the prover: h(X) is a unique polynomial and only its coefficient pieces are
// committed.  BN254's r - 1 is divisible by 3, so a subgroup of order 3 * 2^j exists; the quotient of degree < 5n
// is recovered just as well from 6n points.  This backend therefore works on
//     D = union over c < cosets of  g * zeta^c * <omega_B>,   B = 2^bk * (power of two >= n),  cosets in {1, 3},
// the smallest such set with at least n * (degree - 1) points (3 cosets of 2n for zero_g): 25 % fewer rows for
// evaluate_h and for every resident extended column, and the coefficient -> extended transforms become three 2n-point
// NTTs (54n butterflies at k = 17 against 80n for one 8n-point NTT).  Layout of an extended column: [coset][B];
// the rotation by one row of the 2^k domain is a shift by B/n inside a coset block (powers of two throughout).
// Going back, each coset block is inverse-transformed to P_c(X) = sum_t gamma_c^t h_t(X) (h split into chunks of B
// coefficients, gamma_c = (g zeta^c)^B) and a 3 x 3 inverse Vandermonde recovers the chunks.
// `zg_evaluate_h`, `zg_coeff_to_extended` and `zg_extended_to_coeff` keep halo2's coset at the C ABI.
#pragma once
#include "ctx.cuh"
#include "poly.cuh"

namespace zg {

struct ExtDomain {
  uint32_t k = 0, bk = 0, cosets = 1, rot_scale = 1;
  size_t n = 0, B = 0, N = 0;
  Fr g[3];                   // coset shifts g * zeta^c
  Fr W[9];                   // inverse of V[c][t] = gamma_c^t (cosets == 3)
  Fr omega_B, omega_B_inv;
  Fr* pow_tab = nullptr;     // [cosets][n]  (g_c)^i          (device)
  Fr* inv_tab = nullptr;     // [cosets][B]  (g_c)^-i / B     (device)
  Fr* t_inv = nullptr;       // [cosets][rot_scale]  1 / (X^n - 1) on D (device)
};

// chooses (cosets, B) for a quotient of degree < n * qdeg
void ext_domain_shape(uint32_t k, uint32_t qdeg, ExtDomain& d);
inline size_t ext_domain_table_elems(const ExtDomain& d) { return d.cosets * (d.n + d.B + d.rot_scale) + 8; }
// fills the tables in `mem` (ext_domain_table_elems elements, caller-owned device memory); scan_scratch >= 3 * 2048 Fr,
// fill_buf >= B Fr
int ext_domain_init(zg_ctx* ctx, ExtDomain& d, Fr* mem, Fr* scan_scratch, Fr* fill_buf);
// `batch` polynomials of n coefficients (stride in_stride) -> their values on D (stride out_stride >= N)
// only the coset blocks first, first + step, ... are computed (the others are left untouched): a rank's share of the rows
int ext_from_coeff(zg_ctx* ctx, const ExtDomain& d, const Fr* coeff, size_t in_stride, Fr* out, size_t out_stride, size_t batch,
                   uint32_t first = 0, uint32_t step = 1);
// values on D of a polynomial of degree < N -> its first `keep` coefficients; work holds N elements
int ext_to_coeff(zg_ctx* ctx, const ExtDomain& d, const Fr* ext, Fr* work, size_t keep, Fr* out);
// h[i] /= (X_i^n - 1)
void ext_divide_by_vanishing(const ExtDomain& d, Fr* h, cudaStream_t st, LaunchCounter lc);

}  // namespace zg
