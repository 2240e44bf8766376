// extern "C" surface of libzg_b200.so (declarations + reference citations: include/zg_b200.h).
#include "ctx.cuh"

using namespace zg;

namespace zg {

Fr host_fr_from_u64(uint64_t x) { return fp_from_u64<FrParams>(x); }
Fr host_fr_root_of_unity() {
  // 7^((r-1)/2^28), Montgomery form (halo2curves Fr::ROOT_OF_UNITY)
  Fr raw;
  const uint32_t v[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du,
                         0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
  for (int i = 0; i < 8; i++) raw.v[i] = v[i];
  return fp_to_mont(raw);
}
Fr host_fr_zeta() {
  Fr raw;
  const uint32_t v[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu,
                         0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
  for (int i = 0; i < 8; i++) raw.v[i] = v[i];
  return fp_to_mont(raw);
}
Fr host_omega(uint32_t k) {
  Fr w = host_fr_root_of_unity();
  for (uint32_t i = k; i < 28; i++) w = fp_sqr(w);
  return w;
}

int ws_reserve(zg_ctx* ctx, Workspace& w, size_t bytes) {
  if (w.cap >= bytes) return ZG_OK;
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (w.p) ZG_CUDA(cudaFree(w.p));
  w.p = nullptr;
  w.cap = 0;
  size_t want = bytes + bytes / 8 + 4096;
  cudaError_t e = cudaMalloc(&w.p, want);
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaMalloc workspace: ") + cudaGetErrorString(e);
    return ZG_E_NOMEM;
  }
  w.cap = want;
  return ZG_OK;
}

int get_domain(zg_ctx* ctx, uint32_t logn, const Fr& omega, Domain** out) {
  std::array<uint32_t, 9> key;
  key[0] = logn;
  for (int i = 0; i < 8; i++) key[1 + i] = omega.v[i];
  auto it = ctx->domains.find(key);
  if (it != ctx->domains.end()) {
    *out = &it->second;
    return ZG_OK;
  }
  Domain d;
  size_t n = (size_t)1 << logn;
  ZG_CUDA(cudaMalloc(&d.tw, sizeof(Fr) * n));
  Fr* flat = nullptr;
  cudaError_t e = cudaMalloc(&flat, sizeof(Fr) * (n / 2 + 1));
  if (e == cudaSuccess) {
    e = ntt_build_twiddles(d.tw, flat, omega, logn, ctx->stream);
    ctx->launches += 2;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  if (flat) cudaFree(flat);
  if (e != cudaSuccess) {                 // nothing stays allocated when the table could not be built
    cudaFree(d.tw);
    return ctx->cuda_fail(e, "ntt_build_twiddles");
  }
  auto res = ctx->domains.emplace(key, d);
  *out = &res.first->second;
  return ZG_OK;
}

static Fr fr_from_abi(const zg_fr* p) {
  Fr r;
  memcpy(r.v, p, 32);
  return r;
}

// ---- NTT family on device pointers ---------------------------------------------------------
static int ntt_generic_dev(zg_ctx* ctx, const Fr* in, size_t in_stride, Fr* out, size_t out_stride,
                           uint32_t logn, const Fr& omega, size_t batch, uint32_t n_in, uint32_t n_out,
                           uint32_t flags, const Fr* in_scale, const Fr* out_scale) {
  if (logn < 1 || logn > 28) return ctx->fail(ZG_E_INVALID, "ntt: log_n out of range [1,28]");
  if (batch == 0) return ZG_OK;
  if (batch > 65535) return ctx->fail(ZG_E_INVALID, "ntt: batch > 65535");
  Domain* d;
  int rc = get_domain(ctx, logn, omega, &d);
  if (rc) return rc;
  NttPlan P;
  size_t n = (size_t)1 << logn;
  P.in = in;
  P.out = out;
  P.in_stride = in_stride;
  P.out_stride = out_stride;
  P.tmp = nullptr;
  P.tmp_stride = n;
  if (logn > NTT_MAX_S) {
    rc = ws_reserve(ctx, ctx->ws_ntt, sizeof(Fr) * n * batch);
    if (rc) return rc;
    P.tmp = (Fr*)ctx->ws_ntt.p;
  }
  P.tw = d->tw;
  P.logn = logn;
  P.batch = (uint32_t)batch;
  P.n_in = n_in;
  P.n_out = n_out;
  P.flags = flags;
  for (int i = 0; i < 3; i++) {
    P.in_scale[i] = in_scale ? in_scale[i] : fp_one<FrParams>();
    P.out_scale[i] = out_scale ? out_scale[i] : fp_one<FrParams>();
  }
  cudaError_t e = ntt_run(P, ctx->stream, &ctx->launches);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "ntt_run");
  return ZG_OK;
}

int msm_dev_mixed(zg_ctx* ctx, int basis, const Fr* scalars_dev, size_t stride, size_t n, size_t count, uint32_t other_mask,
                  G1Jac* out_dev) {
  ZG_ENTER(ctx);
  if (!ctx->srs_loaded) return ctx->fail(ZG_E_STATE, "msm: no SRS loaded");
  if (basis < 0 || basis > 1 || !ctx->table[basis].pts) return ctx->fail(ZG_E_STATE, "msm: basis not loaded");
  if (other_mask && (!ctx->table[basis ^ 1].pts || count > 32)) return ctx->fail(ZG_E_STATE, "msm: other basis not loaded");
  const MsmTable& t = ctx->table[basis];
  if (n == 0 || n > t.n) return ctx->fail(ZG_E_INVALID, "msm: n must be in [1, 2^k]");
  if (count == 0) return ZG_OK;
  if ((uint64_t)n * t.W * count >= (1ull << 32) || count > 65535)
    return ctx->fail(ZG_E_INVALID, "msm: batch too large for 32-bit entry indices");
  MsmWorkspaceLayout lay = msm_workspace_layout((uint32_t)n, t.c, t.W, (uint32_t)count);
  int rc = ws_reserve(ctx, ctx->ws_msm, lay.bytes);
  if (rc) return rc;
  cudaError_t e = msm_run(t, scalars_dev, stride, (uint32_t)n, (uint32_t)count, out_dev, ctx->ws_msm.p, lay, ctx->stream,
                          &ctx->launches, &ctx->probe, other_mask ? ctx->table[basis ^ 1].pts : nullptr, other_mask);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "msm_run");
  return ZG_OK;
}

// coefficients (the first n_in of them; the rest read as zero) -> values on halo2's coset zeta * <omega_ext>
int ntt_halo_coset_dev(zg_ctx* ctx, const Fr* coeff, uint32_t n_in, uint32_t ext_k, Fr* out) {
  Fr zeta = host_fr_zeta();
  Fr sc[3] = {fp_one<FrParams>(), zeta, fp_sqr(zeta)};
  return ntt_generic_dev(ctx, coeff, n_in, out, (size_t)1 << ext_k, ext_k, host_omega(ext_k), 1, n_in, 1u << ext_k, NTT_IN_COSET,
                         sc, nullptr);
}

}  // namespace zg

extern "C" {

const char* zg_version(void) { return "zg_b200 0.1 (sm_100a)"; }

int zg_ctx_create(int device, void* stream, zg_ctx** out) {
  if (!out) return ZG_E_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return ZG_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return ZG_E_CUDA;
  zg_ctx* ctx = new zg_ctx();
  ctx->device = device;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return ZG_E_CUDA;
    }
    ctx->own_stream = true;
  }
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  if (cudaMalloc(&ctx->d_msm_out, sizeof(G1Jac) * 64) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->hp, cudaStreamNonBlocking, greatest) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, least) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return ZG_E_CUDA;
  }
  for (int i = 0; i < zg_ctx::N_SIDE; i++) {
    if (cudaStreamCreateWithPriority(&ctx->side[i], cudaStreamNonBlocking, greatest) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_side[i], cudaEventDisableTiming) != cudaSuccess) {
      delete ctx;
      return ZG_E_CUDA;
    }
  }
  *out = ctx;
  return ZG_OK;
}

extern "C" int zg_comm_destroy(zg_ctx* ctx);

void zg_ctx_destroy(zg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) zg_comm_destroy(ctx);
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->d_gather_fr) cudaFree(ctx->d_gather_fr);
  ctx->srs_release();
  for (auto& kv : ctx->domains) cudaFree(kv.second.tw);
  if (ctx->ws_msm.p) cudaFree(ctx->ws_msm.p);
  if (ctx->ws_ntt.p) cudaFree(ctx->ws_ntt.p);
  if (ctx->ws_stage.p) cudaFree(ctx->ws_stage.p);
  if (ctx->d_msm_out) cudaFree(ctx->d_msm_out);
  for (auto e : ctx->probe.ev) cudaEventDestroy(e);
  if (ctx->probe.counts) cudaFreeHost(ctx->probe.counts);
  if (ctx->hp) { cudaStreamSynchronize(ctx->hp); cudaStreamDestroy(ctx->hp); }
  if (ctx->aux) { cudaStreamSynchronize(ctx->aux); cudaStreamDestroy(ctx->aux); }
  for (int i = 0; i < zg_ctx::N_SIDE; i++) {
    if (ctx->side[i]) { cudaStreamSynchronize(ctx->side[i]); cudaStreamDestroy(ctx->side[i]); }
    if (ctx->ev_side[i]) cudaEventDestroy(ctx->ev_side[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* zg_last_error(const zg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t zg_launch_count(const zg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zg_sync(zg_ctx* ctx) {
  ZG_ENTER(ctx);
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}
int zg_dev_alloc(zg_ctx* ctx, size_t bytes, void** out) {
  ZG_ENTER(ctx);
  ZG_CUDA(cudaSetDevice(ctx->device));
  ZG_CUDA(cudaMalloc(out, bytes));
  return ZG_OK;
}
int zg_dev_free(zg_ctx* ctx, void* p) {
  ZG_ENTER(ctx);
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  ZG_CUDA(cudaFree(p));
  return ZG_OK;
}
int zg_h2d(zg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  ZG_ENTER(ctx);
  ZG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}
int zg_d2h(zg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  ZG_ENTER(ctx);
  ZG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

// ---- live kernel timing for bench.py's roofline -----------------------------------------------------
int zg_probe_enable(zg_ctx* ctx, int on) {
  ZG_ENTER(ctx);
  MsmProbe& p = ctx->probe;
  if (on && !p.counts) {
    p.cap = 8192;
    ZG_CUDA(cudaMallocHost(&p.counts, p.cap * sizeof(uint32_t)));
  }
  p.used = 0;
  p.on = on != 0;
  return ZG_OK;
}
int zg_probe_read(zg_ctx* ctx, double* kernel_ms, uint64_t* launches, uint64_t* point_additions) {
  ZG_ENTER(ctx);
  MsmProbe& p = ctx->probe;
  // the entry-count copies were enqueued after the stop events on the same streams: drain everything first
  ZG_CUDA(cudaDeviceSynchronize());
  double ms = 0;
  uint64_t adds = 0;
  for (size_t i = 0; i < p.used; i++) {
    float t = 0;
    ZG_CUDA(cudaEventElapsedTime(&t, p.ev[2 * i], p.ev[2 * i + 1]));
    ms += t;
    adds += p.counts[i];
  }
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = p.used;
  if (point_additions) *point_additions = adds;
  p.used = 0;
  return ZG_OK;
}

// ---- SRS -------------------------------------------------------------------------------------
int zg_srs_load(zg_ctx* ctx, uint32_t k, const zg_g1_affine* g, const zg_g1_affine* g_lagrange) {
  ZG_ENTER(ctx);
  if (k < 1 || k > 26) return ctx->fail(ZG_E_INVALID, "srs_load: k out of range [1,26]");
  ZG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << k;
  const uint32_t c = msm_pick_c(k);
  const uint32_t W = msm_windows(c);
  if ((uint64_t)n * W >= (1ull << 31)) return ctx->fail(ZG_E_INVALID, "srs_load: table index overflow");
  const zg_g1_affine* src[2] = {g, g_lagrange};
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->srs_release();                        // the previous parameters (other contexts that share them keep them alive)
  ctx->srs = new SrsShared();
  for (int b = 0; b < 2; b++) {
    if (!src[b]) continue;
    ZG_CUDA(cudaMalloc(&ctx->srs->base[b], sizeof(G1Affine) * n));
    ctx->base[b] = ctx->srs->base[b];
    ZG_CUDA(cudaMemcpyAsync(ctx->base[b], src[b], sizeof(G1Affine) * n, cudaMemcpyHostToDevice, ctx->stream));
    MsmTable& t = ctx->table[b];
    t.n = (uint32_t)n;
    t.c = c;
    t.W = W;
    cudaError_t e = cudaMalloc(&ctx->srs->table[b], sizeof(G1Affine) * n * W);
    if (e != cudaSuccess) return ctx->fail(ZG_E_NOMEM, "srs_load: window table allocation failed");
    t.pts = ctx->srs->table[b];
    e = msm_precompute_table(ctx->base[b], t.n, c, W, t.pts, ctx->stream);
    ctx->launches += 1;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "msm_precompute_table");
  }
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->srs_k = k;
  ctx->srs_loaded = true;
  return ZG_OK;
}

// `ctx` uses the parameters `from` has loaded (same device): no copy, no second set of window tables
int zg_srs_share(zg_ctx* ctx, zg_ctx* from) {
  ZG_ENTER(ctx);
  if (!from || !from->srs_loaded || !from->srs) return ctx->fail(ZG_E_STATE, "srs_share: the source context has no SRS");
  if (from->device != ctx->device) return ctx->fail(ZG_E_INVALID, "srs_share: contexts on different devices");
  if (from == ctx) return ZG_OK;
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->srs_release();
  from->srs->refs.fetch_add(1);
  ctx->srs = from->srs;
  for (int b = 0; b < 2; b++) { ctx->base[b] = from->base[b]; ctx->table[b] = from->table[b]; }
  ctx->srs_k = from->srs_k;
  ctx->srs_loaded = true;
  return ZG_OK;
}

// ---- MSM -------------------------------------------------------------------------------------
int zg_msm_dev(zg_ctx* ctx, int basis, const zg_fr* scalars_dev, size_t stride, size_t n, size_t count,
               zg_g1* out_dev) {
  return msm_dev_mixed(ctx, basis, (const Fr*)scalars_dev, stride, n, count, 0, (G1Jac*)out_dev);
}

int zg_msm_batch(zg_ctx* ctx, int basis, const zg_fr* const* scalars, size_t n, size_t count, zg_g1* out) {
  ZG_ENTER(ctx);
  if (count == 0) return ZG_OK;
  if (count > 64) return ctx->fail(ZG_E_INVALID, "msm_batch: count > 64");
  int rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * n * count);
  if (rc) return rc;
  Fr* d = (Fr*)ctx->ws_stage.p;
  for (size_t j = 0; j < count; j++)
    ZG_CUDA(cudaMemcpyAsync(d + j * n, scalars[j], sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = zg_msm_dev(ctx, basis, (const zg_fr*)d, n, n, count, (zg_g1*)ctx->d_msm_out);
  if (rc) return rc;
  ZG_CUDA(cudaMemcpyAsync(out, ctx->d_msm_out, sizeof(G1Jac) * count, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

int zg_msm(zg_ctx* ctx, int basis, const zg_fr* scalars, size_t n, zg_g1* out) {
  ZG_ENTER(ctx);
  const zg_fr* arr[1] = {scalars};
  return zg_msm_batch(ctx, basis, arr, n, 1, out);
}

// ---- NTT -------------------------------------------------------------------------------------
int zg_ntt_dev(zg_ctx* ctx, const zg_fr* in_dev, zg_fr* out_dev, uint32_t log_n, const zg_fr* omega,
               size_t batch, size_t stride) {
  ZG_ENTER(ctx);
  Fr w = fr_from_abi(omega);
  uint32_t n = 1u << log_n;
  return ntt_generic_dev(ctx, (const Fr*)in_dev, stride, (Fr*)out_dev, stride, log_n, w, batch, n, n, 0,
                         nullptr, nullptr);
}

int zg_ntt(zg_ctx* ctx, zg_fr* a, uint32_t log_n, const zg_fr* omega) {
  ZG_ENTER(ctx);
  if (log_n < 1 || log_n > 28) return ctx->fail(ZG_E_INVALID, "ntt: log_n out of range [1,28]");
  size_t n = (size_t)1 << log_n;
  int rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * n);
  if (rc) return rc;
  Fr* d = (Fr*)ctx->ws_stage.p;
  ZG_CUDA(cudaMemcpyAsync(d, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = zg_ntt_dev(ctx, (const zg_fr*)d, (zg_fr*)d, log_n, omega, 1, n);
  if (rc) return rc;
  ZG_CUDA(cudaMemcpyAsync(a, d, sizeof(Fr) * n, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

int zg_lagrange_to_coeff_dev(zg_ctx* ctx, const zg_fr* in_dev, zg_fr* out_dev, uint32_t k, size_t batch,
                             size_t stride) {
  ZG_ENTER(ctx);
  if (k < 1 || k > 28) return ctx->fail(ZG_E_INVALID, "lagrange_to_coeff: k out of range");
  Fr winv = fp_inv(host_omega(k));
  Fr ninv = fp_inv(host_fr_from_u64(1ull << k));
  Fr sc[3] = {ninv, ninv, ninv};
  uint32_t n = 1u << k;
  return ntt_generic_dev(ctx, (const Fr*)in_dev, stride, (Fr*)out_dev, stride, k, winv, batch, n, n,
                         NTT_OUT_SCALE, nullptr, sc);
}

int zg_lagrange_to_coeff(zg_ctx* ctx, zg_fr* a, uint32_t k) {
  ZG_ENTER(ctx);
  if (k < 1 || k > 28) return ctx->fail(ZG_E_INVALID, "lagrange_to_coeff: k out of range");
  size_t n = (size_t)1 << k;
  int rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * n);
  if (rc) return rc;
  Fr* d = (Fr*)ctx->ws_stage.p;
  ZG_CUDA(cudaMemcpyAsync(d, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)d, (zg_fr*)d, k, 1, n);
  if (rc) return rc;
  ZG_CUDA(cudaMemcpyAsync(a, d, sizeof(Fr) * n, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

int zg_coeff_to_extended_dev(zg_ctx* ctx, const zg_fr* coeff_dev, size_t in_stride, uint32_t k, uint32_t ext_k,
                             zg_fr* out_dev, size_t out_stride, size_t batch) {
  ZG_ENTER(ctx);
  if (k < 1 || ext_k < k || ext_k > 28) return ctx->fail(ZG_E_INVALID, "coeff_to_extended: bad k/ext_k");
  Fr zeta = host_fr_zeta();
  Fr sc[3] = {fp_one<FrParams>(), zeta, fp_sqr(zeta)};
  return ntt_generic_dev(ctx, (const Fr*)coeff_dev, in_stride, (Fr*)out_dev, out_stride, ext_k,
                         host_omega(ext_k), batch, 1u << k, 1u << ext_k, NTT_IN_COSET, sc, nullptr);
}

int zg_coeff_to_extended(zg_ctx* ctx, const zg_fr* coeff, uint32_t k, uint32_t ext_k, zg_fr* out) {
  ZG_ENTER(ctx);
  if (k < 1 || ext_k < k || ext_k > 28) return ctx->fail(ZG_E_INVALID, "coeff_to_extended: bad k/ext_k");
  size_t n = (size_t)1 << k, ne = (size_t)1 << ext_k;
  int rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * (n + ne));
  if (rc) return rc;
  Fr* din = (Fr*)ctx->ws_stage.p;
  Fr* dout = din + n;
  ZG_CUDA(cudaMemcpyAsync(din, coeff, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = zg_coeff_to_extended_dev(ctx, (const zg_fr*)din, n, k, ext_k, (zg_fr*)dout, ne, 1);
  if (rc) return rc;
  ZG_CUDA(cudaMemcpyAsync(out, dout, sizeof(Fr) * ne, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

int zg_extended_to_coeff_dev(zg_ctx* ctx, const zg_fr* ext_dev, uint32_t k, uint32_t ext_k, size_t keep,
                             zg_fr* out_dev) {
  ZG_ENTER(ctx);
  if (k < 1 || ext_k < k || ext_k > 28) return ctx->fail(ZG_E_INVALID, "extended_to_coeff: bad k/ext_k");
  size_t ne = (size_t)1 << ext_k;
  if (keep == 0 || keep > ne) return ctx->fail(ZG_E_INVALID, "extended_to_coeff: keep out of range");
  Fr winv = fp_inv(host_omega(ext_k));
  Fr ninv = fp_inv(host_fr_from_u64(1ull << ext_k));
  Fr zinv = fp_sqr(host_fr_zeta());  // zeta^-1 = zeta^2
  Fr sc[3] = {ninv, fp_mul(ninv, zinv), fp_mul(ninv, fp_sqr(zinv))};
  return ntt_generic_dev(ctx, (const Fr*)ext_dev, ne, (Fr*)out_dev, ne, ext_k, winv, 1, (uint32_t)ne,
                         (uint32_t)keep, NTT_OUT_SCALE | NTT_OUT_MOD3, nullptr, sc);
}

int zg_extended_to_coeff(zg_ctx* ctx, const zg_fr* ext, uint32_t k, uint32_t ext_k, size_t keep, zg_fr* out) {
  ZG_ENTER(ctx);
  if (k < 1 || ext_k < k || ext_k > 28) return ctx->fail(ZG_E_INVALID, "extended_to_coeff: bad k/ext_k");
  size_t ne = (size_t)1 << ext_k;
  if (keep == 0 || keep > ne) return ctx->fail(ZG_E_INVALID, "extended_to_coeff: keep out of range");
  int rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * 2 * ne);
  if (rc) return rc;
  Fr* din = (Fr*)ctx->ws_stage.p;
  Fr* dout = din + ne;
  ZG_CUDA(cudaMemcpyAsync(din, ext, sizeof(Fr) * ne, cudaMemcpyHostToDevice, ctx->stream));
  rc = zg_extended_to_coeff_dev(ctx, (const zg_fr*)din, k, ext_k, keep, (zg_fr*)dout);
  if (rc) return rc;
  ZG_CUDA(cudaMemcpyAsync(out, dout, sizeof(Fr) * keep, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

}  // extern "C"
