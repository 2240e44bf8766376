// Multi-GPU pieces of the path: one process per GPU, NCCL over NVLink for the one real exchange
// step -- the partial sums / commitments of an MSM round, 96 bytes each.  Everything else of a proof is replicated or
// independent per GPU (an all-to-all of 32-byte field elements costs more than the single-GPU NTT pass at these sizes).
//
//   zg_msm_sharded*      ONE large MSM split by point range (BASELINE configs[3] / [4]): rank g holds bases
//                        [g n/G, (g+1) n/G) and their window table, computes a partial sum; ncclAllGather of G x 96 B
//                        and G - 1 device-side additions (EC addition is not an NCCL reduction, so no all-reduce).
//   spread proof         ONE proof over the ranks (zg_create_proof with ZG_DIST_COLUMNS): every rank runs the same proof on
//                        the same inputs and RNG stream (SPMD) but computes only a share of the two heavy parts --
//                        the commitments of a Fiat-Shamir round by column (column j -> rank j mod G; 96 bytes per
//                        commitment all-gathered, no column crosses NVLink), and the extended forms + quotient numerator
//                        by coset block of the internal extended domain (block c -> rank c mod G; rotations never leave
//                        a block, so only the finished blocks of h, 8 MiB each at k = 17, are all-gathered).
// NCCL is resolved with dlopen at the first use (the library torch has loaded, else the system libnccl.so.2):
// libzg_b200.so itself does not link against it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include "ctx.cuh"

using namespace zg;

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) { api.err = "libnccl.so.2 not found"; return; }
    auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) api.err = std::string("missing NCCL symbol ") + n; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api.err.empty() ? &api : nullptr;
}

int nccl_fail(zg_ctx* ctx, ncclResult_t r, const char* what) {
  NcclApi* a = nccl_api();
  ctx->err = std::string(what) + ": " + (a && a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
  return ZG_E_CUDA;
}

__device__ __forceinline__ G1Xyzz jac_to_xyzz(const G1Jac& p) {
  G1Xyzz r;
  if (fp_is_zero(p.z)) return xyzz_identity();
  r.x = p.x;
  r.y = p.y;
  r.zz = fp_sqr(p.z);
  r.zzz = fp_mul(r.zz, p.z);
  return r;
}

// out[j] = sum_r parts[r * per + j]   (point-range sharding: G partial sums per MSM)
__global__ void k_sum_partials(const G1Jac* __restrict__ parts, uint32_t nranks, uint32_t per, uint32_t count, G1Jac* __restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  G1Xyzz acc = xyzz_identity();
  for (uint32_t r = 0; r < nranks; r++) {
    G1Xyzz p = jac_to_xyzz(parts[(size_t)r * per + j]);
    xyzz_add(acc, p);
  }
  out[j] = xyzz_to_jacobian(acc);
}
// out[j] = gathered[(j % G) * per + j / G]   (column distribution: rank j % G computed commitment j)
__global__ void k_unpack_columns(const G1Jac* __restrict__ gathered, uint32_t nranks, uint32_t per, uint32_t count, G1Jac* __restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  out[j] = gathered[(size_t)(j % nranks) * per + j / nranks];
}

int ensure_gather(zg_ctx* ctx, size_t elems) {
  if (ctx->gather_cap >= elems) return ZG_OK;
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  ctx->d_gather = nullptr;
  ctx->gather_cap = 0;
  cudaError_t e = cudaMalloc(&ctx->d_gather, sizeof(G1Jac) * elems * 2);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "gather buffer");
  ctx->gather_cap = elems;
  return ZG_OK;
}

}  // namespace

namespace zg {

// One Fiat-Shamir round's commitments, results for all `count` columns in ctx->d_msm_out[0..count) on every rank.
int msm_round(zg_ctx* ctx, int basis, const Fr* cols, size_t stride, size_t n, size_t count, uint32_t other_mask) {
  if (ctx->nranks <= 1 || !ctx->comm || !ctx->dist_columns)
    return msm_dev_mixed(ctx, basis, cols, stride, n, count, other_mask, ctx->d_msm_out);
  NcclApi* a = nccl_api();
  if (!a) return ctx->fail(ZG_E_STATE, "NCCL is not available");
  const uint32_t G = (uint32_t)ctx->nranks, rank = (uint32_t)ctx->rank;
  const uint32_t per = ((uint32_t)count + G - 1) / G;
  if (count > 64) return ctx->fail(ZG_E_INVALID, "msm_round: more than 64 commitments in a round");
  int rc = ensure_gather(ctx, (size_t)per * (G + 1));
  if (rc) return rc;
  G1Jac* mine = ctx->d_gather;                     // per slots, then the gathered G * per
  G1Jac* all = ctx->d_gather + per;
  const uint32_t my_count = rank < count ? ((uint32_t)count - rank + G - 1) / G : 0;
  uint32_t my_mask = 0;
  for (uint32_t i = 0; i < my_count; i++)
    if ((other_mask >> (rank + i * G)) & 1u) my_mask |= 1u << i;
  ZG_CUDA(cudaMemsetAsync(mine, 0, sizeof(G1Jac) * per, ctx->stream));
  if (my_count) {
    rc = msm_dev_mixed(ctx, basis, cols + (size_t)rank * stride, stride * G, n, my_count, my_mask, mine);
    if (rc) return rc;
  }
  ncclResult_t r = a->AllGather(mine, all, sizeof(G1Jac) * per, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  k_unpack_columns<<<1, 64, 0, ctx->stream>>>(all, G, per, (uint32_t)count, ctx->d_msm_out);
  ctx->launches++;
  return ZG_OK;
}

int dist_allgather_blocks(zg_ctx* ctx, Fr* col, size_t B, uint32_t nblocks) {
  NcclApi* a = nccl_api();
  if (!a || !ctx->comm) return ctx->fail(ZG_E_STATE, "NCCL is not available");
  const uint32_t G = (uint32_t)ctx->nranks, rank = (uint32_t)ctx->rank;
  const uint32_t per = (nblocks + G - 1) / G;
  const size_t need = (size_t)per * B * (G + 1);
  if (ctx->gather_fr_cap < need) {
    if (ctx->d_gather_fr) cudaFree(ctx->d_gather_fr);
    ctx->d_gather_fr = nullptr;
    ctx->gather_fr_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->d_gather_fr, sizeof(Fr) * need);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "gather buffer");
    ctx->gather_fr_cap = need;
  }
  Fr* mine = ctx->d_gather_fr;                 // per blocks, then the gathered G * per
  Fr* all = mine + (size_t)per * B;
  for (uint32_t i = 0; i < per; i++) {
    const uint32_t c = rank + i * G;
    if (c < nblocks)
      ZG_CUDA(cudaMemcpyAsync(mine + (size_t)i * B, col + (size_t)c * B, sizeof(Fr) * B, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  ncclResult_t r = a->AllGather(mine, all, sizeof(Fr) * per * B, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  for (uint32_t g = 0; g < G; g++)
    for (uint32_t i = 0; i < per; i++) {
      const uint32_t c = g + i * G;
      if (c < nblocks && g != rank)
        ZG_CUDA(cudaMemcpyAsync(col + (size_t)c * B, all + ((size_t)g * per + i) * B, sizeof(Fr) * B, cudaMemcpyDeviceToDevice,
                                ctx->stream));
    }
  return ZG_OK;
}

}  // namespace zg

extern "C" {

int zg_comm_unique_id(uint8_t out[128]) {
  NcclApi* a = nccl_api();
  if (!a || !out) return ZG_E_STATE;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (a->GetUniqueId(&id) != ncclSuccess) return ZG_E_CUDA;
  memcpy(out, &id, 128);
  return ZG_OK;
}

int zg_comm_init(zg_ctx* ctx, int nranks, int rank, const uint8_t unique_id[128]) {
  ZG_ENTER(ctx);
  if (nranks < 1 || rank < 0 || rank >= nranks || !unique_id) return ctx->fail(ZG_E_INVALID, "comm_init: bad rank / size");
  if (ctx->comm) return ctx->fail(ZG_E_STATE, "comm_init: the context already has a communicator");
  NcclApi* a = nccl_api();
  if (!a) return ctx->fail(ZG_E_STATE, "comm_init: NCCL is not available (libnccl.so.2)");
  ncclUniqueId id;
  memcpy(&id, unique_id, 128);
  ncclComm_t comm;
  ncclResult_t r = a->CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclCommInitRank");
  ctx->comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return ZG_OK;
}

int zg_comm_destroy(zg_ctx* ctx) {
  ZG_ENTER(ctx);
  if (ctx->comm) {
    cudaStreamSynchronize(ctx->stream);
    NcclApi* a = nccl_api();
    if (a) a->CommDestroy((ncclComm_t)ctx->comm);
  }
  ctx->comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  return ZG_OK;
}

int zg_ctx_set_distribution(zg_ctx* ctx, int mode) {
  ZG_ENTER(ctx);
  if (mode != ZG_DIST_NONE && mode != ZG_DIST_COLUMNS) return ctx->fail(ZG_E_INVALID, "set_distribution: unknown mode");
  if (mode == ZG_DIST_COLUMNS && !ctx->comm) return ctx->fail(ZG_E_STATE, "set_distribution: no communicator (zg_comm_init)");
  ctx->dist_columns = mode == ZG_DIST_COLUMNS;
  return ZG_OK;
}

// `count` MSMs over this rank's point range (scalars: count columns of n_local, stride apart, on the device); every rank
// receives the `count` totals in out_dev
int zg_msm_sharded_dev(zg_ctx* ctx, int basis, const zg_fr* scalars_dev, size_t stride, size_t n_local, size_t count,
                       zg_g1* out_dev) {
  ZG_ENTER(ctx);
  if (!scalars_dev || !out_dev) return ctx->fail(ZG_E_INVALID, "msm_sharded: null argument");
  if (count == 0) return ZG_OK;
  if (count > 64) return ctx->fail(ZG_E_INVALID, "msm_sharded: more than 64 MSMs");
  if (ctx->nranks <= 1 || !ctx->comm)
    return msm_dev_mixed(ctx, basis, (const Fr*)scalars_dev, stride, n_local, count, 0, (G1Jac*)out_dev);
  NcclApi* a = nccl_api();
  if (!a) return ctx->fail(ZG_E_STATE, "NCCL is not available");
  const uint32_t G = (uint32_t)ctx->nranks, per = (uint32_t)count;
  int rc = ensure_gather(ctx, (size_t)per * (G + 1));
  if (rc) return rc;
  G1Jac* mine = ctx->d_gather;
  G1Jac* all = ctx->d_gather + per;
  rc = msm_dev_mixed(ctx, basis, (const Fr*)scalars_dev, stride, n_local, count, 0, mine);
  if (rc) return rc;
  ncclResult_t r = a->AllGather(mine, all, sizeof(G1Jac) * per, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  k_sum_partials<<<(per + 31) / 32, 32, 0, ctx->stream>>>(all, G, per, per, (G1Jac*)out_dev);
  ctx->launches++;
  return ZG_OK;
}

// host scalars in, host result out (the e2e path of the sharded MSM)
int zg_msm_sharded(zg_ctx* ctx, int basis, const zg_fr* scalars, size_t n_local, zg_g1* out) {
  ZG_ENTER(ctx);
  if (!scalars || !out || n_local == 0) return ctx->fail(ZG_E_INVALID, "msm_sharded: null argument");
  int rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * n_local);
  if (rc) return rc;
  Fr* d = (Fr*)ctx->ws_stage.p;
  ZG_CUDA(cudaMemcpyAsync(d, scalars, sizeof(Fr) * n_local, cudaMemcpyHostToDevice, ctx->stream));
  rc = zg_msm_sharded_dev(ctx, basis, (const zg_fr*)d, n_local, n_local, 1, (zg_g1*)ctx->d_msm_out);
  if (rc) return rc;
  ZG_CUDA(cudaMemcpyAsync(out, ctx->d_msm_out, sizeof(G1Jac), cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

}  // extern "C"
