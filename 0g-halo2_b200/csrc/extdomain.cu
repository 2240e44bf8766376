// Transforms to and from the prover's internal extended domain (extdomain.cuh explains the choice of domain).
#include <cstdlib>
#include "extdomain.cuh"

namespace zg {

namespace {

__device__ __forceinline__ Fr ldx(const Fr* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void stx(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

struct Mat3 { Fr w[9]; };

// out[t*B + i] = sum_c W[t][c] * P[c*B + i]   (t*B + i < keep)
__global__ void k_ext_combine3(const Fr* __restrict__ P, size_t B, Mat3 W, Fr* __restrict__ out, size_t keep) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  Fr p0 = ldx(P + i), p1 = ldx(P + B + i), p2 = ldx(P + 2 * B + i);
#pragma unroll
  for (int t = 0; t < 3; t++) {
    size_t idx = (size_t)t * B + i;
    if (idx >= keep) break;
    Fr v = fp_add(fp_add(fp_mul(W.w[3 * t], p0), fp_mul(W.w[3 * t + 1], p1)), fp_mul(W.w[3 * t + 2], p2));
    stx(out + idx, v);
  }
}

__global__ void k_ext_divide(Fr* __restrict__ h, const Fr* __restrict__ t_inv, uint32_t bk, uint32_t rot_scale, size_t N) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  Fr t = ldx(t_inv + (i >> bk) * rot_scale + (i & (rot_scale - 1)));
  stx(h + i, fp_mul(ldx(h + i), t));
}

}  // namespace

void ext_domain_shape(uint32_t k, uint32_t qdeg, ExtDomain& d) {
  d.k = k;
  d.n = (size_t)1 << k;
  uint32_t b1 = 0, b3 = 0;
  while ((1u << b1) < qdeg) b1++;
  while (3u * (1u << b3) < qdeg) b3++;
  const bool pow2_only = getenv("ZG_EXT_POW2") != nullptr;      // A/B switch: halo2-sized power-of-two domain
  if (!pow2_only && 3u * (1u << b3) < (1u << b1) && k + b3 <= 27) {
    d.cosets = 3;
    d.bk = k + b3;
  } else {
    d.cosets = 1;
    d.bk = k + b1;
  }
  d.B = (size_t)1 << d.bk;
  d.N = d.cosets * d.B;
  d.rot_scale = 1u << (d.bk - k);
}

int ext_domain_init(zg_ctx* ctx, ExtDomain& d, Fr* mem, Fr* scratch, Fr* fill) {
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  d.pow_tab = mem;
  d.inv_tab = d.pow_tab + d.cosets * d.n;
  d.t_inv = d.inv_tab + d.cosets * d.B;
  Fr* one_dev = d.t_inv + d.cosets * d.rot_scale;       // 8 spare elements
  const Fr one = fp_one<FrParams>();
  const Fr gen = host_fr_from_u64(7);                  // Fr::MULTIPLICATIVE_GENERATOR: g^n has no small order
  const Fr zeta = host_fr_zeta();
  d.omega_B = host_omega(d.bk);
  d.omega_B_inv = fp_inv(d.omega_B);
  d.g[0] = gen;
  d.g[1] = fp_mul(gen, zeta);
  d.g[2] = fp_mul(d.g[1], zeta);
  const Fr Binv = fp_inv(host_fr_from_u64((uint64_t)d.B));
  for (uint32_t c = 0; c < d.cosets; c++) {
    ZG_CUDA(cudaMemcpyAsync(one_dev, &one, sizeof(Fr), cudaMemcpyHostToDevice, st));
    fr_fill(fill, d.g[c], d.n, st, lc);
    fr_running_product(fill, one_dev, d.pow_tab + c * d.n, d.n, scratch, st, lc);
    ZG_CUDA(cudaMemcpyAsync(one_dev + 1, &Binv, sizeof(Fr), cudaMemcpyHostToDevice, st));
    fr_fill(fill, fp_inv(d.g[c]), d.B, st, lc);
    fr_running_product(fill, one_dev + 1, d.inv_tab + c * d.B, d.B, scratch, st, lc);
  }
  // 1 / (X^n - 1) at g_c * omega_B^j: g_c^n * (omega_B^n)^j - 1, period rot_scale in j
  std::vector<Fr> t(d.cosets * d.rot_scale);
  const Fr wn = fp_pow_var(d.omega_B, (uint64_t)d.n);
  for (uint32_t c = 0; c < d.cosets; c++) {
    Fr cur = fp_pow_var(d.g[c], (uint64_t)d.n);
    for (uint32_t j = 0; j < d.rot_scale; j++) {
      Fr den = fp_sub(cur, one);
      if (fp_is_zero(den)) return ctx->fail(ZG_E_STATE, "extended domain meets the vanishing set");
      t[c * d.rot_scale + j] = fp_inv(den);
      cur = fp_mul(cur, wn);
    }
  }
  ZG_CUDA(cudaMemcpyAsync(d.t_inv, t.data(), sizeof(Fr) * t.size(), cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaStreamSynchronize(st));                   // `t` is a host temporary
  if (d.cosets == 3) {
    // V[c][t] = gamma_c^t, gamma_c = g_c^B; W = V^-1 by the adjugate (distinct gammas: zeta^(cB) runs over the cube roots)
    Fr gm[3];
    for (int c = 0; c < 3; c++) gm[c] = fp_pow_var(d.g[c], (uint64_t)d.B);
    Fr V[3][3];
    for (int c = 0; c < 3; c++) { V[c][0] = one; V[c][1] = gm[c]; V[c][2] = fp_sqr(gm[c]); }
    auto m2 = [&](int r0, int c0, int r1, int c1) { return fp_sub(fp_mul(V[r0][c0], V[r1][c1]), fp_mul(V[r0][c1], V[r1][c0])); };
    Fr cof[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        const int r0 = (r + 1) % 3, r1 = (r + 2) % 3, c0 = (c + 1) % 3, c1 = (c + 2) % 3;
        cof[r][c] = m2(r0, c0, r1, c1);                  // cyclic indexing carries the sign
      }
    Fr det = fp_add(fp_add(fp_mul(V[0][0], cof[0][0]), fp_mul(V[0][1], cof[0][1])), fp_mul(V[0][2], cof[0][2]));
    if (fp_is_zero(det)) return ctx->fail(ZG_E_STATE, "extended domain: singular coset matrix");
    Fr di = fp_inv(det);
    for (int tt = 0; tt < 3; tt++)
      for (int c = 0; c < 3; c++) d.W[3 * tt + c] = fp_mul(cof[c][tt], di);   // inverse = adjugate^T / det
  }
  return ZG_OK;
}

int ext_from_coeff(zg_ctx* ctx, const ExtDomain& d, const Fr* coeff, size_t in_stride, Fr* out, size_t out_stride, size_t batch,
                   uint32_t first, uint32_t step) {
  if (batch == 0 || first >= d.cosets) return ZG_OK;
  const uint32_t mine = (d.cosets - first + step - 1) / step;
  if (batch * d.cosets > 65535) return ctx->fail(ZG_E_INVALID, "ext_from_coeff: batch too large");
  Domain* dom;
  int rc = get_domain(ctx, d.bk, d.omega_B, &dom);
  if (rc) return rc;
  rc = ws_reserve(ctx, ctx->ws_ntt, sizeof(Fr) * d.B * d.cosets * batch);
  if (rc) return rc;
  NttPlan P;
  P.in = coeff; P.in_stride = in_stride; P.in_coset_stride = 0;
  P.out = out; P.out_stride = out_stride; P.out_coset_stride = d.B;
  P.tmp = (Fr*)ctx->ws_ntt.p; P.tmp_stride = d.B;
  P.tw = dom->tw;
  P.logn = d.bk; P.batch = (uint32_t)batch; P.cosets = mine; P.coset_first = first; P.coset_step = step;
  P.n_in = (uint32_t)d.n; P.n_out = (uint32_t)d.B; P.flags = 0;
  for (int i = 0; i < 3; i++) P.in_scale[i] = P.out_scale[i] = fp_one<FrParams>();
  P.in_table = d.pow_tab; P.in_table_stride = d.n;
  cudaError_t e = ntt_run(P, ctx->stream, &ctx->launches);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "ext_from_coeff");
  return ZG_OK;
}

int ext_to_coeff(zg_ctx* ctx, const ExtDomain& d, const Fr* ext, Fr* work, size_t keep, Fr* out) {
  if (keep == 0 || keep > d.N) return ctx->fail(ZG_E_INVALID, "ext_to_coeff: keep out of range");
  Domain* dom;
  int rc = get_domain(ctx, d.bk, d.omega_B_inv, &dom);
  if (rc) return rc;
  rc = ws_reserve(ctx, ctx->ws_ntt, sizeof(Fr) * d.N);
  if (rc) return rc;
  NttPlan P;
  P.in = ext; P.in_stride = d.N; P.in_coset_stride = d.B;
  P.out = d.cosets == 1 ? out : work; P.out_stride = d.N; P.out_coset_stride = d.B;
  P.tmp = (Fr*)ctx->ws_ntt.p; P.tmp_stride = d.B;
  P.tw = dom->tw;
  P.logn = d.bk; P.batch = 1; P.cosets = d.cosets;
  P.n_in = (uint32_t)d.B; P.n_out = d.cosets == 1 ? (uint32_t)keep : (uint32_t)d.B; P.flags = 0;
  for (int i = 0; i < 3; i++) P.in_scale[i] = P.out_scale[i] = fp_one<FrParams>();
  P.out_table = d.inv_tab; P.out_table_stride = d.B;
  cudaError_t e = ntt_run(P, ctx->stream, &ctx->launches);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "ext_to_coeff");
  if (d.cosets == 3) {
    Mat3 W;
    for (int i = 0; i < 9; i++) W.w[i] = d.W[i];
    k_ext_combine3<<<(unsigned)((d.B + 127) / 128), 128, 0, ctx->stream>>>(work, d.B, W, out, keep);
    ctx->launches++;
  }
  return ZG_OK;
}

void ext_divide_by_vanishing(const ExtDomain& d, Fr* h, cudaStream_t st, LaunchCounter lc) {
  k_ext_divide<<<(unsigned)((d.N + 255) / 256), 256, 0, st>>>(h, d.t_inv, d.bk, d.rot_scale, d.N);
  lc++;
}

}  // namespace zg
