// MSM tail, cooperative form: the warp-shuffle levels of the segmented reduction and the weighted bucket sum with every
// EC addition computed by the four warps of a CTA (msm_coop.cuh: 4 product latencies per addition instead of 14, 3 per
// doubling instead of 9).  Same data flow, same outputs as the kernels of msm_tail.cu -- a CTA here does what ONE warp does
// there: lane l of all four warps holds the same (replicated) values, shuffles are executed by all four warps, products
// are split by warp and exchanged through shared memory, and only warp 0 writes results.
// Used for the small, latency-bound launches of the tail (msm_tail.cu keeps the work-efficient kernels for the large
// ones); ZG_MSM_TAIL_COOP selects.
#define ZG_FP_MUL_NOINLINE 1
#include <cstdlib>
#include "msm.cuh"
#include "msm_coop.cuh"

namespace zg {

namespace {

constexpr int COOP_THREADS = 32 * COOP_ROLES;

// products of one level: [parity][role][limb][lane] (conflict-free 32-bit accesses), double-buffered so that one
// barrier per level suffices (a warp can run at most one level ahead of the slowest)
struct CoopBuf {
  uint32_t w[2][COOP_ROLES][8][32];
};
struct CoopCtx {
  CoopBuf* sb;
  int role, lane, par;
};

__device__ __forceinline__ void coop_put(CoopCtx& c, const Fq& v) {
#pragma unroll
  for (int i = 0; i < 8; i++) c.sb->w[c.par][c.role][i][c.lane] = v.v[i];
}
__device__ __forceinline__ void coop_get_all(const CoopCtx& c, Fq (&o)[COOP_ROLES]) {
#pragma unroll
  for (int r = 0; r < COOP_ROLES; r++)
#pragma unroll
    for (int i = 0; i < 8; i++) o[r].v[i] = c.sb->w[c.par][r][i][c.lane];
}

// acc += b where `active` (all replicated across the four warps); every thread of the CTA must call
__device__ __noinline__ void xyzz_add_coop(G1Xyzz& acc, const G1Xyzz& b, bool active, CoopCtx& c) {
  // nobody in the CTA adds two finite points: no level needed (uniform decision, one barrier)
  if (!__syncthreads_or(coop_add_generic(acc, b, active) ? 1 : 0)) {
    if (active && xyzz_is_identity(acc)) acc = b;
    return;
  }
  CoopAddState s;
  s.a = acc;
  s.b = b;
#pragma unroll   // levels are compile-time: the state stays in registers
  for (int level = 1; level <= COOP_ADD_LEVELS; level++) {
    coop_put(c, coop_add_compute(level, c.role, s));
    __syncthreads();
    Fq o[COOP_ROLES];
    coop_get_all(c, o);
    c.par ^= 1;
    coop_add_absorb(level, s, o);
  }
  acc = coop_add_result(s, active);
}
// p <- 2p (replicated); every thread of the CTA must call
__device__ __noinline__ void xyzz_double_coop(G1Xyzz& p, CoopCtx& c) {
  CoopDblState s;
  s.a = p;
#pragma unroll
  for (int level = 1; level <= COOP_DBL_LEVELS; level++) {
    coop_put(c, coop_dbl_compute(level, c.role, s));
    __syncthreads();
    Fq o[COOP_ROLES];
    coop_get_all(c, o);
    c.par ^= 1;
    coop_dbl_absorb(level, s, o);
  }
  p = coop_dbl_result(s);
}

__device__ __forceinline__ G1Xyzz shfl_down_xyzz(const G1Xyzz& p, int d) {
  G1Xyzz r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], d);
    r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], d);
    r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], d);
    r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], d);
  }
  return r;
}

// lanes hold X_l; lane 0 returns s = sum X_l and t = sum l * X_l (suffix scan, then tree sum): msm_tail.cu::warp_weighted
__device__ void coop_weighted(G1Xyzz x, uint32_t lane, G1Xyzz& s, G1Xyzz& t, CoopCtx& c) {
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    G1Xyzz o = shfl_down_xyzz(x, d);
    xyzz_add_coop(x, o, lane + d < 32, c);
  }
  s = x;
  G1Xyzz y = (lane >= 1) ? x : xyzz_identity();
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    G1Xyzz o = shfl_down_xyzz(y, d);
    xyzz_add_coop(y, o, lane < (uint32_t)d, c);
  }
  t = y;
}
__device__ G1Xyzz coop_sum(G1Xyzz y, uint32_t lane, CoopCtx& c) {
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    G1Xyzz o = shfl_down_xyzz(y, d);
    xyzz_add_coop(y, o, lane < (uint32_t)d, c);
  }
  return y;
}

#define COOP_PROLOGUE                                         \
  __shared__ CoopBuf coop_sb;                                 \
  CoopCtx c;                                                  \
  c.sb = &coop_sb;                                            \
  c.role = threadIdx.x >> 5;                                  \
  c.lane = threadIdx.x & 31;                                  \
  c.par = 0;                                                  \
  const uint32_t lane = threadIdx.x & 31;                     \
  const bool writer = c.role == 0

// ---- warp-shuffle segmented reduction, one group of 32 entries per CTA (msm_tail.cu::msm_warp_reduce_kernel) --------------
__global__ void __launch_bounds__(COOP_THREADS) msm_warp_reduce_coop_kernel(
    const uint32_t* __restrict__ keys, const G1Xyzz* __restrict__ pts, uint32_t n_in, G1Xyzz* __restrict__ buckets,
    uint32_t* __restrict__ pkeys, G1Xyzz* __restrict__ ppts, int final_level) {
  COOP_PROLOGUE;
  const uint32_t g = blockIdx.x;
  const uint32_t e = g * 32 + lane;
  uint32_t key = (e < n_in) ? keys[e] : MSM_INVALID_KEY;
  G1Xyzz acc = (key != MSM_INVALID_KEY) ? pts[e] : xyzz_identity();
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    G1Xyzz other = shfl_down_xyzz(acc, d);
    uint32_t okey = __shfl_down_sync(0xffffffffu, key, d);
    xyzz_add_coop(acc, other, lane + d < 32 && okey == key && key != MSM_INVALID_KEY, c);
  }
  uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  uint32_t last_key = __shfl_sync(0xffffffffu, key, 31);
  if (!writer) return;                                 // (no barrier below this line)
  const bool head = (lane == 0) || (prev != key);
  if (final_level) {
    if (head && key != MSM_INVALID_KEY) buckets[key] = acc;
    return;
  }
  if (key == MSM_INVALID_KEY) {
    if (lane == 0) pkeys[2 * g] = MSM_INVALID_KEY;
    if (lane == 31) pkeys[2 * g + 1] = MSM_INVALID_KEY;
    return;
  }
  if (!head) return;
  const bool touch_end = (last_key == key);
  if (lane == 0) {
    pkeys[2 * g] = key;
    ppts[2 * g] = acc;
    if (touch_end) {
      pkeys[2 * g + 1] = key;
      ppts[2 * g + 1] = xyzz_identity();
    }
  } else if (touch_end) {
    pkeys[2 * g + 1] = key;
    ppts[2 * g + 1] = acc;
  } else {
    buckets[key] = acc;
  }
}

// ---- weighted bucket sum -------------------------------------------------------------------------------------------------
// level 1, one bucket per lane: group g of MSM m = buckets [32g, 32g+32) -> s1[g], t1[g]  (msm_bucket_l1_warp_kernel)
__global__ void __launch_bounds__(COOP_THREADS) msm_bucket_l1_coop_kernel(const G1Xyzz* __restrict__ buckets, uint32_t NB,
                                                                           uint32_t n1, G1Xyzz* __restrict__ s1,
                                                                           G1Xyzz* __restrict__ t1) {
  COOP_PROLOGUE;
  const uint32_t g = blockIdx.x, m = blockIdx.y;
  const uint32_t b = g * 32 + lane;
  G1Xyzz x = (b < NB) ? buckets[(size_t)m * NB + b] : xyzz_identity();
  G1Xyzz s, t;
  coop_weighted(x, lane, s, t, c);
  if (writer && lane == 0) {
    s1[(size_t)m * n1 + g] = s;
    t1[(size_t)m * n1 + g] = t;
  }
}

// level 2: CTA (w, m, task): task 0 folds s1[32w..32w+32) -> (S2, T2), task 1 sums t1[32w..32w+32) -> U  (msm_bucket_l2_kernel)
__global__ void __launch_bounds__(COOP_THREADS) msm_bucket_l2_coop_kernel(const G1Xyzz* __restrict__ s1,
                                                                           const G1Xyzz* __restrict__ t1, uint32_t n1,
                                                                           G1Xyzz* __restrict__ l2out) {
  COOP_PROLOGUE;
  const uint32_t w = blockIdx.x, m = blockIdx.y, task = blockIdx.z, nw = gridDim.x;
  const uint32_t i = w * 32 + lane;
  G1Xyzz* o = l2out + (size_t)m * 3 * nw;
  if (task == 0) {
    G1Xyzz x = (i < n1) ? s1[(size_t)m * n1 + i] : xyzz_identity();
    G1Xyzz s, t;
    coop_weighted(x, lane, s, t, c);
    if (writer && lane == 0) {
      o[w] = s;
      o[nw + w] = t;
    }
  } else {
    G1Xyzz tt = (i < n1) ? t1[(size_t)m * n1 + i] : xyzz_identity();
    G1Xyzz u = coop_sum(tt, lane, c);
    if (writer && lane == 0) o[2 * nw + w] = u;
  }
}

// finish: one CTA per MSM folds the nw <= 32 (S2, T2, U) triples:
//   W = sum U + 32 * ( sum T2 + 32 * t3 ),  result = W + S   (msm_finish_kernel)
__global__ void __launch_bounds__(COOP_THREADS) msm_finish_coop_kernel(const G1Xyzz* __restrict__ l2out, uint32_t nw,
                                                                        G1Jac* __restrict__ out) {
  COOP_PROLOGUE;
  const uint32_t m = blockIdx.x;
  const G1Xyzz* o = l2out + (size_t)m * 3 * nw;
  G1Xyzz S, t3;
  coop_weighted((lane < nw) ? o[lane] : xyzz_identity(), lane, S, t3, c);
  G1Xyzz T = coop_sum((lane < nw) ? o[nw + lane] : xyzz_identity(), lane, c);
  G1Xyzz U = coop_sum((lane < nw) ? o[2 * nw + lane] : xyzz_identity(), lane, c);
  // lane 0 of every warp holds S, t3, T, U; the remaining chain runs on lane 0's values (all lanes execute it)
  G1Xyzz r = t3;
#pragma unroll 1
  for (int i = 0; i < 5; i++) xyzz_double_coop(r, c);
  xyzz_add_coop(r, T, true, c);
#pragma unroll 1
  for (int i = 0; i < 5; i++) xyzz_double_coop(r, c);
  xyzz_add_coop(r, U, true, c);
  xyzz_add_coop(r, S, true, c);
  if (writer && lane == 0) out[m] = xyzz_to_jacobian(r);
}

}  // namespace

// ---- launchers (called from msm.cu / msm_tail.cu) ------------------------------------------------------------------------------
bool msm_tail_coop_enabled() {
  static const int v = [] {
    const char* e = getenv("ZG_MSM_TAIL_COOP");
    return e ? atoi(e) : ZG_MSM_TAIL_COOP_DEFAULT;
  }();
  return v != 0;
}
void msm_tail_warp_level_coop(const uint32_t* keys, const G1Xyzz* pts, uint32_t slots, G1Xyzz* buckets, uint32_t* pkeys_out,
                              G1Xyzz* ppts_out, uint32_t nwarps, int fin, cudaStream_t st) {
  msm_warp_reduce_coop_kernel<<<nwarps, COOP_THREADS, 0, st>>>(keys, pts, slots, buckets, pkeys_out, ppts_out, fin);
}
void msm_tail_l1_coop(const G1Xyzz* buckets, uint32_t NB, uint32_t M, uint32_t n1, G1Xyzz* s1, G1Xyzz* t1, cudaStream_t st) {
  msm_bucket_l1_coop_kernel<<<dim3(n1, M), COOP_THREADS, 0, st>>>(buckets, NB, n1, s1, t1);
}
void msm_tail_l2_finish_coop(const G1Xyzz* s1, const G1Xyzz* t1, uint32_t n1, uint32_t M, G1Xyzz* l2out, G1Jac* out,
                             cudaStream_t st) {
  const uint32_t nw = (n1 + 31) / 32;
  msm_bucket_l2_coop_kernel<<<dim3(nw, M, 2), COOP_THREADS, 0, st>>>(s1, t1, n1, l2out);
  msm_finish_coop_kernel<<<M, COOP_THREADS, 0, st>>>(l2out, nw, out);
}

}  // namespace zg
