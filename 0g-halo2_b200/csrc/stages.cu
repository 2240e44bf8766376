// Per-function entry points of the prover stages (host pointers in, host pointers out).
//
// zg_create_proof (prover.cu) runs a whole proof with every column resident on the device; a halo2_proofs
// fork that wants to replace ONE upstream function at a time binds these instead.  Each wraps exactly the
// kernels create_proof uses for that step, so the parity tests of tests/test_gpu_stages.py cover the same
// code the proof path runs.  Upstream functions (halo2_proofs v2023_04_20, un-vendored,
// /root/reference/Cargo.toml:21-25) are named at every entry point; all are reached from create_proof,
// call site /root/reference/src/wnn.rs:242-259.
#include <algorithm>
#include "ctx.cuh"
#include "lookup.cuh"
#include "poly.cuh"

using namespace zg;

namespace {

struct Carve {
  uint8_t* p;
  size_t off = 0;
  template <class T>
  T* take(size_t count) {
    T* r = (T*)(p + off);
    off += (count * sizeof(T) + 255) & ~(size_t)255;
    return r;
  }
};
inline size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

// plonk::lookup::prover::permute_expression_pair (without the blinding rows): a, s = compressed input / table
// expressions over the first `usable` rows.
int zg_lookup_permute(zg_ctx* ctx, const zg_fr* a, const zg_fr* s, size_t usable, zg_fr* a_perm, zg_fr* s_perm) {
  ZG_ENTER(ctx);
  if (!a || !s || !a_perm || !s_perm) return ctx->fail(ZG_E_INVALID, "lookup_permute: null argument");
  if (usable == 0) return ZG_OK;
  if (usable >= (1u << 28)) return ctx->fail(ZG_E_INVALID, "lookup_permute: too many rows");
  const uint32_t n = (uint32_t)usable;
  const size_t col = pad(sizeof(Fr) * usable);
  int rc = ws_reserve(ctx, ctx->ws_stage, 4 * col + pad(lookup_workspace_bytes(n)) + 256);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  Carve c{ctx->ws_stage.p};
  Fr* da = c.take<Fr>(usable);
  Fr* ds = c.take<Fr>(usable);
  Fr* dpa = c.take<Fr>(usable);
  Fr* dps = c.take<Fr>(usable);
  uint32_t* flags = c.take<uint32_t>(2);
  uint8_t* ws = c.take<uint8_t>(lookup_workspace_bytes(n));
  ZG_CUDA(cudaMemcpyAsync(da, a, sizeof(Fr) * usable, cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaMemcpyAsync(ds, s, sizeof(Fr) * usable, cudaMemcpyHostToDevice, st));
  LookupTable tab = lookup_workspace_table(ws, n);
  // as in zg_create_proof: sort on the top 48 bits first; the device flags distinct keys left out of order (values that
  // agree in their top 48 bits, e.g. small integers) and the sort is redone on all 256 bits
  uint32_t hflags[2] = {0, 0};
  for (int attempt = 0; attempt < 2; attempt++) {
    ZG_CUDA(cudaMemsetAsync(flags, 0, 8, st));
    if (lookup_sort_table(ds, n, tab, ws, flags, /*full_sort=*/attempt == 1, st, lc))
      return ctx->cuda_fail(cudaGetLastError(), "lookup_sort_table");
    if (lookup_permute(da, n, tab, dpa, dps, ws, flags + 1, st, lc)) return ctx->cuda_fail(cudaGetLastError(), "lookup_permute");
    ZG_CUDA(cudaMemcpyAsync(hflags, flags, 8, cudaMemcpyDeviceToHost, st));
    ZG_CUDA(cudaStreamSynchronize(st));
    if (!hflags[0]) break;
  }
  ZG_CUDA(cudaMemcpyAsync(a_perm, dpa, sizeof(Fr) * usable, cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaMemcpyAsync(s_perm, dps, sizeof(Fr) * usable, cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaStreamSynchronize(st));
  if (hflags[1]) return ctx->fail(ZG_E_SYNTH, "lookup_permute: input value not in table (ConstraintSystemFailure)");
  return ZG_OK;
}

// The grand-product step shared by plonk::permutation::prover::commit and plonk::lookup::prover::commit_product:
// z[0] = 1, z[i] = z[i-1] * num[i-1] / den[i-1] for i < len (denominators inverted with one batch inversion).
int zg_grand_product(zg_ctx* ctx, const zg_fr* num, const zg_fr* den, size_t len, zg_fr* z) {
  ZG_ENTER(ctx);
  if (!num || !den || !z) return ctx->fail(ZG_E_INVALID, "grand_product: null argument");
  if (len == 0) return ZG_OK;
  const size_t col = pad(sizeof(Fr) * len);
  int rc = ws_reserve(ctx, ctx->ws_stage, 3 * col + pad(sizeof(Fr) * 4096) + 512);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  Carve c{ctx->ws_stage.p};
  Fr* dn = c.take<Fr>(len);
  Fr* dd = c.take<Fr>(len);
  Fr* dz = c.take<Fr>(len);
  Fr* scratch = c.take<Fr>(4096);
  ZG_CUDA(cudaMemcpyAsync(dn, num, sizeof(Fr) * len, cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaMemcpyAsync(dd, den, sizeof(Fr) * len, cudaMemcpyHostToDevice, st));
  fr_batch_invert(dd, len, st, lc);
  fr_mul_vec(dn, dd, dd, len, st, lc);
  const uint32_t nout = (uint32_t)len;
  fr_running_product_batch(dd, len, dz, len, &nout, 1, scratch, st, lc);
  ZG_CUDA(cudaMemcpyAsync(z, dz, sizeof(Fr) * len, cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaStreamSynchronize(st));
  return cudaGetLastError() == cudaSuccess ? ZG_OK : ctx->cuda_fail(cudaGetLastError(), "grand_product");
}

// ff::BatchInvert::batch_invert (zeros stay zero), in place
int zg_batch_invert(zg_ctx* ctx, zg_fr* a, size_t n) {
  ZG_ENTER(ctx);
  if (!a) return ctx->fail(ZG_E_INVALID, "batch_invert: null argument");
  if (n == 0) return ZG_OK;
  int rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * n);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  Fr* d = (Fr*)ctx->ws_stage.p;
  ZG_CUDA(cudaMemcpyAsync(d, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, st));
  fr_batch_invert(d, n, st, lc);
  ZG_CUDA(cudaMemcpyAsync(a, d, sizeof(Fr) * n, cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaStreamSynchronize(st));
  return ZG_OK;
}

// arithmetic::eval_polynomial for `count` polynomials of n coefficients at one point x
int zg_eval_poly_batch(zg_ctx* ctx, const zg_fr* const* polys, size_t n, size_t count, const zg_fr* x, zg_fr* out) {
  ZG_ENTER(ctx);
  if (!polys || !x || !out) return ctx->fail(ZG_E_INVALID, "eval_poly_batch: null argument");
  if (count == 0) return ZG_OK;
  if (n == 0 || count > 4096) return ctx->fail(ZG_E_INVALID, "eval_poly_batch: n == 0 or count > 4096");
  const size_t col = pad(sizeof(Fr) * n);
  const size_t per = std::max<size_t>(64, (n + 4095) / 4096);   // partial sums per polynomial (fr_eval_many)
  int rc = ws_reserve(ctx, ctx->ws_stage, count * col + pad(sizeof(Fr) * count * per) + pad(count * 16) + pad(sizeof(Fr) * count) + 1024);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  Carve c{ctx->ws_stage.p};
  std::vector<const Fr*> ptrs(count);
  for (size_t j = 0; j < count; j++) {
    Fr* d = c.take<Fr>(n);
    ptrs[j] = d;
    ZG_CUDA(cudaMemcpyAsync(d, polys[j], sizeof(Fr) * n, cudaMemcpyHostToDevice, st));
  }
  Fr* scratch = c.take<Fr>(count * per);
  const Fr** dptrs = c.take<const Fr*>(count);
  uint32_t* pidx = c.take<uint32_t>(count);
  Fr* dout = c.take<Fr>(count);
  Fr* dx = c.take<Fr>(1);
  ZG_CUDA(cudaMemcpyAsync(dptrs, ptrs.data(), count * sizeof(Fr*), cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaMemsetAsync(pidx, 0, count * 4, st));
  ZG_CUDA(cudaMemcpyAsync(dx, x, sizeof(Fr), cudaMemcpyHostToDevice, st));
  fr_eval_many(dptrs, pidx, dx, (uint32_t)count, n, dout, scratch, st, lc);
  ZG_CUDA(cudaMemcpyAsync(out, dout, sizeof(Fr) * count, cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaStreamSynchronize(st));   // also keeps `ptrs` alive until the pointer upload has completed
  return ZG_OK;
}

// arithmetic::kate_division: q(X) = (a(X) - a(z)) / (X - z); a has n coefficients, q has n - 1
int zg_kate_division(zg_ctx* ctx, const zg_fr* a, size_t n, const zg_fr* z, zg_fr* q) {
  ZG_ENTER(ctx);
  if (!a || !z || !q) return ctx->fail(ZG_E_INVALID, "kate_division: null argument");
  if (n < 2) return ZG_OK;
  const size_t col = pad(sizeof(Fr) * n);
  int rc = ws_reserve(ctx, ctx->ws_stage, 2 * col + pad(sizeof(Fr) * 4 * 2048) + 256);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  Carve c{ctx->ws_stage.p};
  Fr* da = c.take<Fr>(n);
  Fr* dq = c.take<Fr>(n);
  Fr* scratch = c.take<Fr>(4 * 2048);
  Fr zz;
  memcpy(zz.v, z, 32);
  ZG_CUDA(cudaMemcpyAsync(da, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, st));
  fr_kate_division(da, n, zz, dq, scratch, st, lc);
  ZG_CUDA(cudaMemcpyAsync(q, dq, sizeof(Fr) * (n - 1), cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaStreamSynchronize(st));
  return ZG_OK;
}

}  // extern "C"
