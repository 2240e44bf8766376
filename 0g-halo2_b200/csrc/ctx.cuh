// Per-GPU context: stream, SRS window tables, twiddle-table cache, grow-only workspace.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <array>
#include <atomic>
#include "../../include/zg_b200.h"
#include "msm.cuh"
#include "ntt.cuh"

namespace zg {

struct Domain {       // twiddle table for one (log_n, omega)
  Fr* tw = nullptr;   // stage-major, n-1 entries
};

// SRS bases and their window tables on one device: shared by the contexts that proved they want the same parameters
// (zg_srs_share), freed with the last of them
struct SrsShared {
  std::atomic<int> refs{1};
  G1Affine* base[2] = {nullptr, nullptr};
  G1Affine* table[2] = {nullptr, nullptr};
  ~SrsShared() {
    for (int b = 0; b < 2; b++) {
      if (base[b]) cudaFree(base[b]);
      if (table[b]) cudaFree(table[b]);
    }
  }
};

struct Workspace {
  uint8_t* p = nullptr;
  size_t cap = 0;
};

// host-side field helpers (portable path of field.cuh)
Fr host_fr_from_u64(uint64_t x);
Fr host_fr_root_of_unity();      // order 2^28
Fr host_fr_zeta();
Fr host_omega(uint32_t k);       // ROOT_OF_UNITY ^ (2^(28-k))

}  // namespace zg

struct zg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // create_proof runs its critical path (MSM -> transcript) on `hp` (highest priority) and the transforms nobody waits
  // for until the quotient stage on `aux` (lowest priority); both are forked from / joined to `stream` with events
  cudaStream_t hp = nullptr, aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // independent latency-bound chains of one stage (the per-lookup sort + permutation) run side by side on these
  static constexpr int N_SIDE = 4;
  cudaStream_t side[N_SIDE] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_side[N_SIDE] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
  uint64_t launches = 0;

  // SRS
  uint32_t srs_k = 0;
  bool srs_loaded = false;
  zg::SrsShared* srs = nullptr;               // owner of the device memory behind base[] / table[]
  zg::G1Affine* base[2] = {nullptr, nullptr};
  zg::MsmTable table[2];
  void srs_release() {
    if (srs && srs->refs.fetch_sub(1) == 1) delete srs;
    srs = nullptr;
    for (int b = 0; b < 2; b++) { base[b] = nullptr; table[b] = zg::MsmTable(); }
    srs_loaded = false;
  }

  // twiddle tables keyed by (log_n, omega limbs)
  std::map<std::array<uint32_t, 9>, zg::Domain> domains;

  // multi-GPU (dist.cu): NCCL communicator (opaque here), this rank, and whether zg_create_proof spreads a round's
  // commitments over the ranks
  void* comm = nullptr;
  int nranks = 1, rank = 0;
  bool dist_columns = false;
  zg::G1Jac* d_gather = nullptr;
  size_t gather_cap = 0;
  zg::Fr* d_gather_fr = nullptr;       // staging of dist_allgather_blocks
  size_t gather_fr_cap = 0;

  zg::Workspace ws_msm, ws_ntt, ws_stage;
  zg::G1Jac* d_msm_out = nullptr;  // small result staging (64 results)
  zg::MsmProbe probe;              // zg_probe_enable / zg_probe_read

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return ZG_E_CUDA;
  }
};

namespace zg {
int ws_reserve(zg_ctx* ctx, Workspace& w, size_t bytes);
int get_domain(zg_ctx* ctx, uint32_t logn, const Fr& omega, Domain** out);
// zg_msm_dev with MSM m on the OTHER basis when bit m of other_mask is set (count <= 32)
int msm_dev_mixed(zg_ctx* ctx, int basis, const Fr* scalars_dev, size_t stride, size_t n, size_t count, uint32_t other_mask,
                  G1Jac* out_dev);
// the commitments of one round into ctx->d_msm_out[0..count) -- locally, or spread over the ranks of ctx->comm (dist.cu)
int msm_round(zg_ctx* ctx, int basis, const Fr* cols, size_t stride, size_t n, size_t count, uint32_t other_mask);
// col holds nblocks blocks of B elements; block c was computed by rank c mod G: all-gather so that every rank holds all of them
int dist_allgather_blocks(zg_ctx* ctx, Fr* col, size_t B, uint32_t nblocks);
}  // namespace zg

// CUDA's current device is per host thread: every entry point selects the context's device first
#define ZG_ENTER(ctx)                                              \
  do {                                                             \
    if (!(ctx)) return ZG_E_INVALID;                               \
    cudaError_t e__ = cudaSetDevice((ctx)->device);                \
    if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, "cudaSetDevice"); \
  } while (0)

#define ZG_CUDA(call)                                             \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return ctx->cuda_fail(e__, #call);    \
  } while (0)
