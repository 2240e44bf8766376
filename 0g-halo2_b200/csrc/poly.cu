// Column / polynomial kernels of the prover: blinding-scalar reduction, batch inversion, grand
// products (parallel prefix products), Kate division (parallel first-order recurrence), Horner
// evaluation at the opening points and the linear folds of the GWC opening.
//
// Replaces the host loops of halo2_proofs v2023_04_20 (un-vendored; pinned by
// /root/reference/Cargo.toml:21-25) in plonk/permutation/prover.rs (commit), plonk/lookup/prover.rs
// (commit_product), plonk/vanishing/prover.rs, arithmetic::{eval_polynomial, kate_division} and
// poly/kzg/multiopen/gwc/prover.rs, all reached from create_proof
// (/root/reference/src/wnn.rs:242-259).  The upstream scans are serial; here they are
// chunk -> one-CTA block scan -> chunk, and bit-exact because field multiplication is associative.
#include "poly.cuh"

namespace zg {

namespace {

constexpr int EW_THREADS = 256;
constexpr uint32_t SCAN_MAX_CHUNKS = 2048;  // one 1024-thread CTA scans two chunk summaries per thread

__device__ __forceinline__ Fr ldf(const Fr* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void stf(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

inline uint32_t blocks_for(size_t n, int threads = EW_THREADS) { return (uint32_t)((n + threads - 1) / threads); }

// ---- elementwise -----------------------------------------------------------------------------
__global__ void k_from_u512(const uint64_t* __restrict__ w, Fr* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr d0, d1, r2, r3;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint64_t a = w[8 * i + j], b = w[8 * i + 4 + j];
    d0.v[2 * j] = (uint32_t)a; d0.v[2 * j + 1] = (uint32_t)(a >> 32);
    d1.v[2 * j] = (uint32_t)b; d1.v[2 * j + 1] = (uint32_t)(b >> 32);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) { r2.v[j] = FrParams::r2(j); r3.v[j] = FrParams::r3(j); }
  // Fr::from_u512: d0 * R^2 + d1 * R^3.  The unreduced 256-bit digit must be the operand whose limbs
  // drive the outer CIOS loop (second argument) so every intermediate stays below 2p.
  stf(out + i, fp_add(fp_mul(r2, d0), fp_mul(r3, d1)));
}
__global__ void k_fill(Fr* a, Fr v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(a + i, v);
}
__global__ void k_mul_add_scalar(const Fr* __restrict__ a, Fr s, const Fr* __restrict__ b, Fr* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr r = fp_mul(ldf(a + i), s);
  if (b) r = fp_add(r, ldf(b + i));
  stf(out + i, r);
}
__global__ void k_mul_periodic(Fr* a, const Fr* __restrict__ t, uint32_t period, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(a + i, fp_mul(ldf(a + i), ldf(t + (i % period))));
}
__global__ void k_one_minus_sum(const Fr* __restrict__ a, const Fr* __restrict__ b, Fr* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(out + i, fp_sub(fp_sub(fp_one<FrParams>(), ldf(a + i)), ldf(b + i)));
}
__global__ void k_scatter_rows(Fr* a, const uint32_t* __restrict__ idx, const Fr* __restrict__ v, uint32_t m) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) stf(a + idx[j], ldf(v + j));
}

// ---- batch inversion --------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ Fp<P> ldp(const Fp<P>* p) {
  Fp<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class P>
__device__ __forceinline__ void stp(Fp<P>* p, const Fp<P>& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
constexpr int BI_CHUNK = 16;
template <class P>
__global__ void __launch_bounds__(128) k_batch_invert(Fp<P>* a, size_t n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t b = t * BI_CHUNK;
  if (b >= n) return;
  int len = (int)((n - b < (size_t)BI_CHUNK) ? (n - b) : BI_CHUNK);
  Fp<P> pre[BI_CHUNK];
  Fp<P> acc = fp_one<P>();
  for (int i = 0; i < len; i++) {
    pre[i] = acc;
    Fp<P> x = ldp(a + b + i);
    if (!fp_is_zero(x)) acc = fp_mul(acc, x);
  }
  Fp<P> inv = fp_inv(acc);
  for (int i = len - 1; i >= 0; i--) {
    Fp<P> x = ldp(a + b + i);
    if (fp_is_zero(x)) continue;
    stp(a + b + i, fp_mul(inv, pre[i]));
    inv = fp_mul(inv, x);
  }
}

// Two-level form for long vectors.  One Fermat inversion (~380 products) per 16-element chunk makes the kernel above cost
// ~27 products per element (0.44 ms for the 6 x 2^17 grand-product denominators of a k = 17 proof); here every thread
// multiplies a strided chunk of BI2_CHUNK elements (coalesced: element i of thread t is a[i * T + t]), the T chunk products
// are inverted by the kernel above, and a second pass turns them into the element inverses: 4 products per element.
constexpr int BI2_CHUNK = 32;
template <class P>
__global__ void __launch_bounds__(128) k_bi2_products(const Fp<P>* __restrict__ a, size_t n, size_t T, Fp<P>* __restrict__ totals) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  Fp<P> acc = fp_one<P>();
  for (int i = 0; i < BI2_CHUNK; i++) {
    size_t idx = (size_t)i * T + t;
    if (idx >= n) break;
    Fp<P> x = ldp(a + idx);
    if (!fp_is_zero(x)) acc = fp_mul(acc, x);
  }
  stp(totals + t, acc);
}
template <class P>
__global__ void __launch_bounds__(128) k_bi2_apply(Fp<P>* __restrict__ a, size_t n, size_t T, const Fp<P>* __restrict__ totals_inv) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  Fp<P> pre[BI2_CHUNK];
  Fp<P> acc = fp_one<P>();
  int len = 0;
  for (int i = 0; i < BI2_CHUNK; i++) {
    size_t idx = (size_t)i * T + t;
    if (idx >= n) break;
    pre[i] = acc;
    Fp<P> x = ldp(a + idx);
    if (!fp_is_zero(x)) acc = fp_mul(acc, x);
    len = i + 1;
  }
  Fp<P> inv = ldp(totals_inv + t);
  for (int i = len - 1; i >= 0; i--) {
    size_t idx = (size_t)i * T + t;
    Fp<P> x = ldp(a + idx);
    if (fp_is_zero(x)) continue;
    stp(a + idx, fp_mul(inv, pre[i]));
    inv = fp_mul(inv, x);
  }
}

// ---- scans ---------------------------------------------------------------------------------------
struct ScanGeom {
  uint32_t chunks, len;
};
inline ScanGeom scan_geom(size_t n) {
  ScanGeom g;
  uint32_t len = 32;
  while ((n + len - 1) / len > SCAN_MAX_CHUNKS) len *= 2;
  g.len = len;
  g.chunks = (uint32_t)((n + len - 1) / len);
  return g;
}

// phase 1 of the running product over f[0 .. m): P[c] = prod of chunk c
__global__ void k_rp_chunk(const Fr* __restrict__ f, size_t m, uint32_t len, Fr* __restrict__ P) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  size_t b = (size_t)c * len;
  if (b >= m) return;
  size_t e = b + len < m ? b + len : m;
  Fr acc = ldf(f + b);
  for (size_t i = b + 1; i < e; i++) acc = fp_mul(acc, ldf(f + i));
  stf(P + c, acc);
}
// phase 2: one CTA, E[c] = start * prod_{c' < c} P[c']   (Hillis-Steele over pairs)
__global__ void __launch_bounds__(1024) k_rp_block_scan(const Fr* __restrict__ P, uint32_t chunks, const Fr* __restrict__ start,
                                                        Fr* __restrict__ E) {
  __shared__ Fr sm[1024];
  const uint32_t t = threadIdx.x;
  Fr one = fp_one<FrParams>();
  Fr a0 = (2 * t < chunks) ? ldf(P + 2 * t) : one;
  Fr a1 = (2 * t + 1 < chunks) ? ldf(P + 2 * t + 1) : one;
  Fr x = fp_mul(a0, a1);
  sm[t] = x;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    Fr o = (t >= d) ? sm[t - d] : one;
    __syncthreads();
    if (t >= d) {
      x = fp_mul(x, o);
      sm[t] = x;
    }
    __syncthreads();
  }
  Fr excl = (t > 0) ? sm[t - 1] : one;
  Fr s = ldf(start);
  Fr e0 = fp_mul(s, excl);
  if (2 * t < chunks) stf(E + 2 * t, e0);
  if (2 * t + 1 < chunks) stf(E + 2 * t + 1, fp_mul(e0, a0));
}
// phase 3: z[b] = E[c]; z[i+1] = z[i] * f[i] inside the chunk
__global__ void k_rp_apply(const Fr* __restrict__ f, const Fr* __restrict__ E, uint32_t len, Fr* __restrict__ z, size_t n_out) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  size_t b = (size_t)c * len;
  if (b >= n_out) return;
  size_t e = b + len < n_out ? b + len : n_out;
  Fr acc = ldf(E + c);
  stf(z + b, acc);
  for (size_t i = b + 1; i < e; i++) {
    acc = fp_mul(acc, ldf(f + i - 1));
    stf(z + i, acc);
  }
}

// Kate division as the forward recurrence b[t] = A[t] + zz * b[t-1], A[t] = a[n-1-t], q[n-2-t] = b[t]
__device__ __forceinline__ void k_kd_chunk_body(const Fr* __restrict__ a, size_t n, const Fr& zz, uint32_t len, Fr* __restrict__ Lc) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  size_t m = n - 1;
  size_t b = (size_t)c * len;
  if (b >= m) return;
  size_t e = b + len < m ? b + len : m;
  Fr acc = fp_zero<FrParams>();
  for (size_t t = b; t < e; t++) acc = fp_add(ldf(a + (n - 1 - t)), fp_mul(zz, acc));
  stf(Lc + c, acc);
}
__global__ void k_kd_chunk(const Fr* __restrict__ a, size_t n, Fr zz, uint32_t len, Fr* __restrict__ Lc) {
  k_kd_chunk_body(a, n, zz, len, Lc);
}
// carry[c] = b[c*len - 1] = sum_{c' < c} M^(c-1-c') L[c'], M = zz^len
__device__ __forceinline__ void k_kd_block_scan_body(const Fr* __restrict__ Lc, uint32_t chunks, const Fr& M, Fr* __restrict__ carry) {
  __shared__ Fr sm[1024];
  const uint32_t t = threadIdx.x;
  Fr zero = fp_zero<FrParams>();
  Fr l0 = (2 * t < chunks) ? ldf(Lc + 2 * t) : zero;
  Fr l1 = (2 * t + 1 < chunks) ? ldf(Lc + 2 * t + 1) : zero;
  // pair summary: value after both chunks with zero carry-in = l1 + M*l0 ; pair multiplier M^2
  Fr x = fp_add(l1, fp_mul(M, l0));
  Fr mult = fp_sqr(M);
  sm[t] = x;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    Fr o = (t >= d) ? sm[t - d] : zero;
    __syncthreads();
    if (t >= d) {
      x = fp_add(x, fp_mul(mult, o));
      sm[t] = x;
    }
    mult = fp_sqr(mult);
    __syncthreads();
  }
  Fr excl = (t > 0) ? sm[t - 1] : zero;   // carry into chunk 2t
  if (2 * t < chunks) stf(carry + 2 * t, excl);
  if (2 * t + 1 < chunks) stf(carry + 2 * t + 1, fp_add(l0, fp_mul(M, excl)));
}
__global__ void __launch_bounds__(1024) k_kd_block_scan(const Fr* __restrict__ Lc, uint32_t chunks, Fr M, Fr* __restrict__ carry) {
  k_kd_block_scan_body(Lc, chunks, M, carry);
}
__device__ __forceinline__ void k_kd_apply_body(const Fr* __restrict__ a, size_t n, const Fr& zz, uint32_t len,
                                                const Fr* __restrict__ carry, Fr* __restrict__ q) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  size_t m = n - 1;
  size_t b = (size_t)c * len;
  if (b >= m) return;
  size_t e = b + len < m ? b + len : m;
  Fr acc = ldf(carry + c);
  for (size_t t = b; t < e; t++) {
    acc = fp_add(ldf(a + (n - 1 - t)), fp_mul(zz, acc));
    stf(q + (m - 1 - t), acc);
  }
}
__global__ void k_kd_apply(const Fr* __restrict__ a, size_t n, Fr zz, uint32_t len, const Fr* __restrict__ carry, Fr* __restrict__ q) {
  k_kd_apply_body(a, n, zz, len, carry, q);
}

// ---- evaluation at points ---------------------------------------------------------------------------
constexpr int EV_PER_THREAD = 64;   // Horner run per thread; its x^(64 t) offset is (x^64)^t: ~1.35 products per coefficient
constexpr int EV_THREADS = 256;
// grid (blocks, count); partial[j * gridDim.x + blockIdx.x]
__global__ void __launch_bounds__(EV_THREADS) k_eval_partial(const Fr* const* __restrict__ polys, const uint32_t* __restrict__ pidx,
                                                             const Fr* __restrict__ points, size_t n, Fr* __restrict__ partial) {
  __shared__ Fr sm[EV_THREADS];
  const uint32_t j = blockIdx.y;
  const Fr* p = polys[j];
  const Fr x = ldf(points + pidx[j]);
  size_t t = (size_t)blockIdx.x * EV_THREADS + threadIdx.x;
  size_t b = t * EV_PER_THREAD;
  Fr acc = fp_zero<FrParams>();
  if (b < n) {
    size_t e = b + EV_PER_THREAD < n ? b + EV_PER_THREAD : n;
    for (size_t i = e; i-- > b;) acc = fp_add(fp_mul(acc, x), ldf(p + i));
    Fr xs = x;
#pragma unroll 1
    for (int sq = 1; sq < EV_PER_THREAD; sq <<= 1) xs = fp_sqr(xs);       // x^EV_PER_THREAD
    acc = fp_mul(acc, fp_pow_var(xs, (uint64_t)t));
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int d = EV_THREADS / 2; d >= 1; d >>= 1) {
    if ((int)threadIdx.x < d) sm[threadIdx.x] = fp_add(sm[threadIdx.x], sm[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) stf(partial + (size_t)j * gridDim.x + blockIdx.x, sm[0]);
}
__global__ void k_eval_final(const Fr* __restrict__ partial, uint32_t per, uint32_t count, Fr* __restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  Fr acc = fp_zero<FrParams>();
  for (uint32_t b = 0; b < per; b++) acc = fp_add(acc, ldf(partial + (size_t)j * per + b));
  stf(out + j, acc);
}

__global__ void k_linear_combination(const Fr* const* __restrict__ polys, const Fr* __restrict__ coeff, uint32_t count, size_t n,
                                     Fr sub0, Fr* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr acc = fp_zero<FrParams>();
  for (uint32_t j = 0; j < count; j++) acc = fp_add(acc, fp_mul(ldf(polys[j] + i), ldf(coeff + j)));
  if (i == 0) acc = fp_sub(acc, sub0);
  stf(out + i, acc);
}

__global__ void k_sigma_values(const uint32_t* __restrict__ mapping, const Fr* __restrict__ delta_pow, Fr omega, uint32_t m, size_t n,
                               Fr* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)m * n) return;
  uint32_t c2 = mapping[2 * i], r2 = mapping[2 * i + 1];
  stf(out + i, fp_mul(ldf(delta_pow + c2), fp_pow_var(omega, (uint64_t)r2)));
}

__global__ void k_mul_vec(const Fr* __restrict__ a, const Fr* __restrict__ b, Fr* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(out + i, fp_mul(ldf(a + i), ldf(b + i)));
}
__global__ void k_perm_fraction(const Fr* const* __restrict__ vals, const Fr* const* __restrict__ sigmas, uint32_t count,
                                const Fr* __restrict__ wpow, Fr beta, Fr gamma, Fr delta_start, Fr delta, Fr* __restrict__ num,
                                Fr* __restrict__ den, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr dw = fp_mul(fp_mul(delta_start, ldf(wpow + i)), beta);   // delta^j * w^i * beta
  Fr nu = fp_one<FrParams>(), de = fp_one<FrParams>();
  for (uint32_t j = 0; j < count; j++) {
    Fr v = ldf(vals[j] + i);
    de = fp_mul(de, fp_add(fp_add(fp_mul(beta, ldf(sigmas[j] + i)), gamma), v));
    nu = fp_mul(nu, fp_add(fp_add(dw, gamma), v));
    dw = fp_mul(dw, delta);
  }
  stf(num + i, nu);
  stf(den + i, de);
}
__global__ void k_lookup_fraction(const Fr* __restrict__ ci, const Fr* __restrict__ ct, const Fr* __restrict__ pa,
                                  const Fr* __restrict__ ps, Fr beta, Fr gamma, Fr* __restrict__ num, Fr* __restrict__ den, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  stf(num + i, fp_mul(fp_add(ldf(ci + i), beta), fp_add(ldf(ct + i), gamma)));
  stf(den + i, fp_mul(fp_add(ldf(pa + i), beta), fp_add(ldf(ps + i), gamma)));
}

struct RpBatch {
  uint32_t n_out[16];
};
__global__ void k_rpb_chunk(const Fr* __restrict__ f, size_t f_stride, RpBatch B, uint32_t len, Fr* __restrict__ P) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t col = blockIdx.y;
  const size_t m = B.n_out[col];
  size_t b = (size_t)c * len;
  if (b >= m) return;
  size_t e = b + len < m ? b + len : m;
  const Fr* fc = f + col * f_stride;
  Fr acc = ldf(fc + b);
  for (size_t i = b + 1; i < e; i++) acc = fp_mul(acc, ldf(fc + i));
  stf(P + (size_t)col * 2 * SCAN_MAX_CHUNKS + c, acc);
}
__global__ void __launch_bounds__(1024) k_rpb_block_scan(Fr* __restrict__ PE, RpBatch B, uint32_t len) {
  __shared__ Fr sm[1024];
  const uint32_t t = threadIdx.x;
  const uint32_t col = blockIdx.x;
  const uint32_t chunks = (B.n_out[col] + len - 1) / len;
  const Fr* P = PE + (size_t)col * 2 * SCAN_MAX_CHUNKS;
  Fr* E = PE + (size_t)col * 2 * SCAN_MAX_CHUNKS + SCAN_MAX_CHUNKS;
  Fr one = fp_one<FrParams>();
  Fr a0 = (2 * t < chunks) ? ldf(P + 2 * t) : one;
  Fr a1 = (2 * t + 1 < chunks) ? ldf(P + 2 * t + 1) : one;
  Fr x = fp_mul(a0, a1);
  sm[t] = x;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    Fr o = (t >= d) ? sm[t - d] : one;
    __syncthreads();
    if (t >= d) {
      x = fp_mul(x, o);
      sm[t] = x;
    }
    __syncthreads();
  }
  Fr e0 = (t > 0) ? sm[t - 1] : one;
  if (2 * t < chunks) stf(E + 2 * t, e0);
  if (2 * t + 1 < chunks) stf(E + 2 * t + 1, fp_mul(e0, a0));
}
__global__ void k_rpb_apply(const Fr* __restrict__ f, size_t f_stride, const Fr* __restrict__ PE, RpBatch B, uint32_t len,
                            Fr* __restrict__ z, size_t z_stride) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t col = blockIdx.y;
  const size_t n_out = B.n_out[col];
  size_t b = (size_t)c * len;
  if (b >= n_out) return;
  size_t e = b + len < n_out ? b + len : n_out;
  const Fr* fc = f + col * f_stride;
  Fr* zc = z + col * z_stride;
  Fr acc = ldf(PE + (size_t)col * 2 * SCAN_MAX_CHUNKS + SCAN_MAX_CHUNKS + c);
  stf(zc + b, acc);
  for (size_t i = b + 1; i < e; i++) {
    acc = fp_mul(acc, ldf(fc + i - 1));
    stf(zc + i, acc);
  }
}
__global__ void k_scale_by_dev(Fr* a, const Fr* __restrict__ s, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(a + i, fp_mul(ldf(a + i), ldf(s)));
}

}  // namespace

void fr_running_product_batch(const Fr* f, size_t f_stride, Fr* z, size_t z_stride, const uint32_t* n_out, uint32_t count,
                               Fr* scratch, cudaStream_t st, LaunchCounter lc) {
  if (!count) return;
  RpBatch B;
  uint32_t mx = 0;
  for (uint32_t i = 0; i < count; i++) { B.n_out[i] = n_out[i]; mx = n_out[i] > mx ? n_out[i] : mx; }
  ScanGeom g = scan_geom(mx);
  k_rpb_chunk<<<dim3(blocks_for(g.chunks, 128), count), 128, 0, st>>>(f, f_stride, B, g.len, scratch);
  lc++;
  k_rpb_block_scan<<<count, 1024, 0, st>>>(scratch, B, g.len);
  lc++;
  k_rpb_apply<<<dim3(blocks_for(g.chunks, 128), count), 128, 0, st>>>(f, f_stride, scratch, B, g.len, z, z_stride);
  lc++;
}
void fr_scale_by_dev(Fr* a, const Fr* scalar_dev, size_t n, cudaStream_t st, LaunchCounter lc) {
  k_scale_by_dev<<<blocks_for(n), EW_THREADS, 0, st>>>(a, scalar_dev, n);
  lc++;
}
void fr_mul_vec(const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc) {
  k_mul_vec<<<blocks_for(n), EW_THREADS, 0, st>>>(a, b, out, n);
  lc++;
}
void perm_fraction(const Fr* const* vals, const Fr* const* sigmas, uint32_t count, const Fr* wpow, const Fr& beta, const Fr& gamma,
                   const Fr& delta_start, const Fr& delta, Fr* num, Fr* den, size_t n, cudaStream_t st, LaunchCounter lc) {
  k_perm_fraction<<<blocks_for(n), EW_THREADS, 0, st>>>(vals, sigmas, count, wpow, beta, gamma, delta_start, delta, num, den, n);
  lc++;
}
void lookup_fraction(const Fr* ci, const Fr* ct, const Fr* pa, const Fr* ps, const Fr& beta, const Fr& gamma, Fr* num, Fr* den,
                     size_t n, cudaStream_t st, LaunchCounter lc) {
  k_lookup_fraction<<<blocks_for(n), EW_THREADS, 0, st>>>(ci, ct, pa, ps, beta, gamma, num, den, n);
  lc++;
}

// ---- host wrappers ---------------------------------------------------------------------------------
void fr_from_u512(const uint64_t* words, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc) {
  if (!n) return;
  k_from_u512<<<blocks_for(n), EW_THREADS, 0, st>>>(words, out, n);
  lc++;
}
void fr_fill(Fr* a, const Fr& v, size_t n, cudaStream_t st, LaunchCounter lc) {
  if (!n) return;
  k_fill<<<blocks_for(n), EW_THREADS, 0, st>>>(a, v, n);
  lc++;
}
void fr_mul_add_scalar(const Fr* a, const Fr& s, const Fr* b, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc) {
  if (!n) return;
  k_mul_add_scalar<<<blocks_for(n), EW_THREADS, 0, st>>>(a, s, b, out, n);
  lc++;
}
void fr_mul_periodic(Fr* a, const Fr* t_dev, uint32_t period, size_t n, cudaStream_t st, LaunchCounter lc) {
  if (!n) return;
  k_mul_periodic<<<blocks_for(n), EW_THREADS, 0, st>>>(a, t_dev, period, n);
  lc++;
}
void fr_one_minus_sum(const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t st, LaunchCounter lc) {
  k_one_minus_sum<<<blocks_for(n), EW_THREADS, 0, st>>>(a, b, out, n);
  lc++;
}
void fr_scatter_rows(Fr* a, const uint32_t* idx_dev, const Fr* v_dev, uint32_t m, cudaStream_t st, LaunchCounter lc) {
  if (!m) return;
  k_scatter_rows<<<blocks_for(m, 64), 64, 0, st>>>(a, idx_dev, v_dev, m);
  lc++;
}
template <class P>
static void batch_invert_any(Fp<P>* a, size_t n, cudaStream_t st, LaunchCounter lc, Fp<P>* scratch) {
  if (!n) return;
  if (scratch && n >= (size_t)1 << 14) {        // two-level: chunk products -> their inverses -> element inverses
    const size_t T = (n + BI2_CHUNK - 1) / BI2_CHUNK;
    k_bi2_products<P><<<blocks_for(T, 128), 128, 0, st>>>(a, n, T, scratch);
    lc++;
    k_batch_invert<P><<<blocks_for((T + BI_CHUNK - 1) / BI_CHUNK, 128), 128, 0, st>>>(scratch, T);
    lc++;
    k_bi2_apply<P><<<blocks_for(T, 128), 128, 0, st>>>(a, n, T, scratch);
    lc++;
    return;
  }
  size_t threads = (n + BI_CHUNK - 1) / BI_CHUNK;
  k_batch_invert<P><<<blocks_for(threads, 128), 128, 0, st>>>(a, n);
  lc++;
}
void fr_batch_invert(Fr* a, size_t n, cudaStream_t st, LaunchCounter lc, Fr* scratch) { batch_invert_any<FrParams>(a, n, st, lc, scratch); }
void fr_running_product(const Fr* f, const Fr* start_dev, Fr* z, size_t n_out, Fr* scratch, cudaStream_t st, LaunchCounter lc) {
  if (!n_out) return;
  // chunks are laid over z (n_out entries); chunk c's product covers f[c*len .. (c+1)*len)
  ScanGeom g = scan_geom(n_out);
  Fr* P = scratch;
  Fr* E = scratch + SCAN_MAX_CHUNKS;
  k_rp_chunk<<<blocks_for(g.chunks, 128), 128, 0, st>>>(f, n_out, g.len, P);
  lc++;
  k_rp_block_scan<<<1, 1024, 0, st>>>(P, g.chunks, start_dev, E);
  lc++;
  k_rp_apply<<<blocks_for(g.chunks, 128), 128, 0, st>>>(f, E, g.len, z, n_out);
  lc++;
}
void fr_kate_division(const Fr* a, size_t n, const Fr& z, Fr* q, Fr* scratch, cudaStream_t st, LaunchCounter lc) {
  if (n < 2) return;
  ScanGeom g = scan_geom(n - 1);
  Fr* Lc = scratch;
  Fr* carry = scratch + SCAN_MAX_CHUNKS;
  Fr M = fp_pow_var(z, (uint64_t)g.len);
  k_kd_chunk<<<blocks_for(g.chunks, 128), 128, 0, st>>>(a, n, z, g.len, Lc);
  lc++;
  k_kd_block_scan<<<1, 1024, 0, st>>>(Lc, g.chunks, M, carry);
  lc++;
  k_kd_apply<<<blocks_for(g.chunks, 128), 128, 0, st>>>(a, n, z, g.len, carry, q);
  lc++;
}
void fr_eval_many(const Fr* const* polys_dev, const uint32_t* point_idx_dev, const Fr* points_dev, uint32_t count, size_t n,
                  Fr* out_dev, Fr* scratch, cudaStream_t st, LaunchCounter lc) {
  if (!count) return;
  uint32_t per = (uint32_t)((n + (size_t)EV_THREADS * EV_PER_THREAD - 1) / ((size_t)EV_THREADS * EV_PER_THREAD));
  k_eval_partial<<<dim3(per, count), EV_THREADS, 0, st>>>(polys_dev, point_idx_dev, points_dev, n, scratch);
  lc++;
  k_eval_final<<<blocks_for(count, 64), 64, 0, st>>>(scratch, per, count, out_dev);
  lc++;
}
// ---- batched GWC witnesses: sets are grid.y ------------------------------------------------------------
__global__ void k_gwc_fold(const Fr* const* __restrict__ polys, const Fr* __restrict__ coeff, GwcBatch B, size_t n, Fr* __restrict__ fold) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = blockIdx.y;
  if (i >= n) return;
  Fr acc = fp_zero<FrParams>();
  for (uint32_t j = B.off[g]; j < B.off[g + 1]; j++) acc = fp_add(acc, fp_mul(ldf(polys[j] + i), ldf(coeff + j)));
  if (i == 0) acc = fp_sub(acc, B.sub0[g]);
  stf(fold + (size_t)g * n + i, acc);
}
__global__ void k_gwc_kd_chunk(const Fr* __restrict__ fold, size_t n, GwcBatch B, uint32_t len, Fr* __restrict__ scratch) {
  const uint32_t g = blockIdx.y;
  k_kd_chunk_body(fold + (size_t)g * n, n, B.z[g], len, scratch + (size_t)g * 2 * SCAN_MAX_CHUNKS);
}
__global__ void __launch_bounds__(1024) k_gwc_kd_block_scan(GwcBatch B, uint32_t chunks, Fr* __restrict__ scratch) {
  const uint32_t g = blockIdx.x;
  Fr* Lc = scratch + (size_t)g * 2 * SCAN_MAX_CHUNKS;
  k_kd_block_scan_body(Lc, chunks, B.M[g], Lc + SCAN_MAX_CHUNKS);
}
__global__ void k_gwc_kd_apply(const Fr* __restrict__ fold, size_t n, GwcBatch B, uint32_t len, const Fr* __restrict__ scratch,
                               Fr* __restrict__ q) {
  const uint32_t g = blockIdx.y;
  k_kd_apply_body(fold + (size_t)g * n, n, B.z[g], len, scratch + (size_t)g * 2 * SCAN_MAX_CHUNKS + SCAN_MAX_CHUNKS, q + (size_t)g * n);
}

void fr_gwc_witness_batch(const Fr* const* polys_dev, const Fr* coeff_dev, GwcBatch B, size_t n, Fr* fold, Fr* q, Fr* scratch,
                          cudaStream_t st, LaunchCounter lc) {
  if (!B.nsets || n < 2) return;
  ScanGeom g = scan_geom(n - 1);
  for (uint32_t s = 0; s < B.nsets; s++) B.M[s] = fp_pow_var(B.z[s], (uint64_t)g.len);
  k_gwc_fold<<<dim3(blocks_for(n), B.nsets), EW_THREADS, 0, st>>>(polys_dev, coeff_dev, B, n, fold);
  lc++;
  k_gwc_kd_chunk<<<dim3(blocks_for(g.chunks, 128), B.nsets), 128, 0, st>>>(fold, n, B, g.len, scratch);
  lc++;
  k_gwc_kd_block_scan<<<B.nsets, 1024, 0, st>>>(B, g.chunks, scratch);
  lc++;
  k_gwc_kd_apply<<<dim3(blocks_for(g.chunks, 128), B.nsets), 128, 0, st>>>(fold, n, B, g.len, scratch, q);
  lc++;
}

void fr_linear_combination(const Fr* const* polys_dev, const Fr* coeff_dev, uint32_t count, size_t n, const Fr& sub0, Fr* out,
                           cudaStream_t st, LaunchCounter lc) {
  k_linear_combination<<<blocks_for(n), EW_THREADS, 0, st>>>(polys_dev, coeff_dev, count, n, sub0, out);
  lc++;
}
void fr_sigma_values(const uint32_t* mapping_dev, const Fr* delta_pow_dev, const Fr& omega, uint32_t m, size_t n, Fr* out,
                     cudaStream_t st, LaunchCounter lc) {
  k_sigma_values<<<blocks_for((size_t)m * n), EW_THREADS, 0, st>>>(mapping_dev, delta_pow_dev, omega, m, n, out);
  lc++;
}

}  // namespace zg
