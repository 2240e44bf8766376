// Polynomial / column kernels used by the prover orchestration (poly.cu, expr.cu, lookup.cu).
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace zg {

struct LaunchCounter {
  uint64_t* n;
  void operator++(int) { if (n) ++*n; }
};

// ---- elementwise (poly.cu) -------------------------------------------------------------------
void fr_from_u512(const uint64_t* words, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc);
void fr_fill(Fr* a, const Fr& v, size_t n, cudaStream_t st, LaunchCounter lc);
// out[i] = a[i] * s + b[i]   (b may be null -> a[i]*s)
void fr_mul_add_scalar(const Fr* a, const Fr& s, const Fr* b, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc);
// a[i] *= t[i % period]   (divide_by_vanishing_poly with precomputed inverses)
void fr_mul_periodic(Fr* a, const Fr* t_dev, uint32_t period, size_t n, cudaStream_t st, LaunchCounter lc);
// out[i] = 1 - a[i] - b[i]
void fr_one_minus_sum(const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc);
// a[idx[j]] = v[j] for j < m (tiny scatter of blinding rows etc.)
void fr_scatter_rows(Fr* a, const uint32_t* idx_dev, const Fr* v_dev, uint32_t m, cudaStream_t st, LaunchCounter lc);

// batch inversion in place (zeros stay zero), Montgomery trick per thread chunk
// scratch (optional): ceil(n / 32) elements; enables the two-level form for n >= 2^14
void fr_batch_invert(Fr* a, size_t n, cudaStream_t st, LaunchCounter lc, Fr* scratch = nullptr);
// z[0] = *start_dev, z[i] = z[i-1] * f[i-1] for i < n_out  (grand products); scratch >= 3 * 2048 Fr
void fr_running_product(const Fr* f, const Fr* start_dev, Fr* z, size_t n_out, Fr* scratch, cudaStream_t st,
                        LaunchCounter lc);
// `count` (<= 16) independent running products with start 1: column c reads f + c*f_stride and writes
// n_out[c] entries of z + c*z_stride; scratch >= count * 4096 Fr
void fr_running_product_batch(const Fr* f, size_t f_stride, Fr* z, size_t z_stride, const uint32_t* n_out, uint32_t count,
                               Fr* scratch, cudaStream_t st, LaunchCounter lc);
// a[i] *= *scalar_dev
void fr_scale_by_dev(Fr* a, const Fr* scalar_dev, size_t n, cudaStream_t st, LaunchCounter lc);
// kate_division: q has n-1 coefficients, q[i-1] = a[i] + z*q[i]; a[0] is first replaced by a[0]-eval
// (done by the caller).  scratch >= 4 * 2048 Fr
void fr_kate_division(const Fr* a, size_t n, const Fr& z, Fr* q, Fr* scratch, cudaStream_t st, LaunchCounter lc);

// polynomial evaluation: out[j] = polys[j](points[point_idx[j]]), all polys have n coefficients.
// `polys_dev` is a device array of device pointers.  scratch >= count * 64 Fr
void fr_eval_many(const Fr* const* polys_dev, const uint32_t* point_idx_dev, const Fr* points_dev, uint32_t count,
                  size_t n, Fr* out_dev, Fr* scratch, cudaStream_t st, LaunchCounter lc);
// out[i] = sum_j coeff[j] * polys[j][i], then out[0] -= sub0   (GWC fold by powers of v; h_poly from pieces)
void fr_linear_combination(const Fr* const* polys_dev, const Fr* coeff_dev, uint32_t count, size_t n, const Fr& sub0,
                           Fr* out, cudaStream_t st, LaunchCounter lc);
// ProverGWC witnesses for up to 8 opening points in four launches: set g folds polys[off[g] .. off[g+1]) with the
// matching coefficients (powers of v), subtracts sub0[g] (the folded evaluation) from the constant term and divides
// by (X - z[g]).  fold and q hold nsets columns of n; scratch >= nsets * 4096 Fr.  M is filled in by the callee.
struct GwcBatch {
  uint32_t nsets;
  uint32_t off[9];
  Fr sub0[8], z[8], M[8];
};
void fr_gwc_witness_batch(const Fr* const* polys_dev, const Fr* coeff_dev, GwcBatch B, size_t n, Fr* fold, Fr* q, Fr* scratch,
                          cudaStream_t st, LaunchCounter lc);
// out[i] = a[i] * b[i]
void fr_mul_vec(const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc);
// permutation argument, one column set: num[i] = prod_j (v_j[i] + delta_start*delta^j * beta * w^i + gamma),
// den[i] = prod_j (v_j[i] + beta * sigma_j[i] + gamma); vals / sigmas are device arrays of column pointers
void perm_fraction(const Fr* const* vals, const Fr* const* sigmas, uint32_t count, const Fr* wpow, const Fr& beta,
                   const Fr& gamma, const Fr& delta_start, const Fr& delta, Fr* num, Fr* den, size_t n, cudaStream_t st,
                   LaunchCounter lc);
// lookup argument: num[i] = (ci+beta)(ct+gamma), den[i] = (pa+beta)(ps+gamma)
void lookup_fraction(const Fr* ci, const Fr* ct, const Fr* pa, const Fr* ps, const Fr& beta, const Fr& gamma, Fr* num,
                     Fr* den, size_t n, cudaStream_t st, LaunchCounter lc);
// permutation sigma values: out[c][r] = delta_pow[c'] * omega^{r'} with mapping[c][r] = (c', r')
void fr_sigma_values(const uint32_t* mapping_dev, const Fr* delta_pow_dev, const Fr& omega, uint32_t m, size_t n,
                     Fr* out, cudaStream_t st, LaunchCounter lc);

}  // namespace zg
