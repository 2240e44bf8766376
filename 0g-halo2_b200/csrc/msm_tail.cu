// MSM tail kernels: folding of chunk-boundary partial sums and the weighted bucket sum.
//
// Same algorithm as described at the top of msm.cu (which keeps the digit extraction, the counting sort
// and the level-0 accumulation).  These kernels run with one warp per SM or less, i.e. they are bound by
// the latency of a chain of EC additions, so this translation unit outlines the field multiplier
// (ZG_FP_MUL_NOINLINE, see field.cuh) to keep the instruction footprint inside the instruction cache.
#define ZG_FP_MUL_NOINLINE 1
#include "msm.cuh"

namespace zg {

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

// ---- serial segmented reduction over a list of (key, XYZZ partial sum) entries -------------------------
// (level 0, over the packed table-index entries, is msm_accumulate_kernel in msm.cu.)
// Thread t owns entries [t*K, t*K+K).  Runs strictly inside the chunk go straight to their bucket
// (nobody else holds that key); the first and last runs go to partial slots 2t, 2t+1.
__global__ void __launch_bounds__(128) msm_serial_reduce_kernel(
    const uint32_t* __restrict__ keys, const G1Xyzz* __restrict__ pts_in, uint32_t L, uint32_t K,
    G1Xyzz* __restrict__ buckets, uint32_t* __restrict__ pkeys, G1Xyzz* __restrict__ ppts, uint32_t nthreads) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  const uint64_t start64 = (uint64_t)t * K;
  uint32_t cur = start64 < L ? keys[start64] : MSM_INVALID_KEY;
  if (cur == MSM_INVALID_KEY) {   // past the end, or the list ends in invalid slots
    pkeys[2 * t] = MSM_INVALID_KEY;
    pkeys[2 * t + 1] = MSM_INVALID_KEY;
    return;
  }
  const uint32_t start = (uint32_t)start64;
  const uint32_t end = (start64 + K < L) ? start + K : L;
  G1Xyzz acc = xyzz_identity();
  uint32_t nruns = 0;
  for (uint32_t e = start; e < end; e++) {
    uint32_t k = keys[e];
    if (k == MSM_INVALID_KEY) break;
    if (k != cur) {
      if (nruns == 0) {
        pkeys[2 * t] = cur;
        ppts[2 * t] = acc;
      } else {
        buckets[cur] = acc;
      }
      nruns++;
      cur = k;
      acc = xyzz_identity();
    }
    G1Xyzz p = pts_in[e];
    xyzz_add(acc, p);
  }
  if (nruns == 0) {
    pkeys[2 * t] = cur;
    ppts[2 * t] = acc;
    pkeys[2 * t + 1] = cur;  // same key, identity value: keeps runs contiguous downstream
    ppts[2 * t + 1] = xyzz_identity();
  } else {
    pkeys[2 * t + 1] = cur;
    ppts[2 * t + 1] = acc;
  }
}

// ---- warp-shuffle segmented reduction (small lists) ------------------------------------
// One entry per lane.  After 5 shuffle steps the head lane of every run holds the run's sum
// inside this warp.  Runs touching the warp's edges go to slots 2g / 2g+1 of the next level;
// interior runs (and every run when `final_level`) are written to their bucket.
__global__ void __launch_bounds__(128) msm_warp_reduce_kernel(
    const uint32_t* __restrict__ keys, const G1Xyzz* __restrict__ pts, uint32_t n_in,
    G1Xyzz* __restrict__ buckets, uint32_t* __restrict__ pkeys, G1Xyzz* __restrict__ ppts,
    uint32_t nwarps, int final_level) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (g >= nwarps) return;  // whole warps only: blockDim is a multiple of 32
  const uint32_t e = g * 32 + lane;
  uint32_t key = (e < n_in) ? keys[e] : MSM_INVALID_KEY;
  G1Xyzz acc = (key != MSM_INVALID_KEY) ? pts[e] : xyzz_identity();
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    G1Xyzz other = shfl_down_xyzz(acc, d);
    uint32_t okey = __shfl_down_sync(0xffffffffu, key, d);
    if (lane + d < 32 && okey == key && key != MSM_INVALID_KEY) xyzz_add(acc, other);
  }
  uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  uint32_t last_key = __shfl_sync(0xffffffffu, key, 31);
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

// ---- weighted bucket sum ---------------------------------------------------------------
// lanes hold X_l; returns in lane 0: s = sum X_l and t = sum l * X_l (suffix scan + tree sum).
__device__ __forceinline__ void warp_weighted(G1Xyzz x, uint32_t lane, G1Xyzz& s, G1Xyzz& t) {
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    G1Xyzz o = shfl_down_xyzz(x, d);
    if (lane + d < 32) xyzz_add(x, o);
  }
  s = x;  // lane 0: total
  G1Xyzz y = (lane >= 1) ? x : xyzz_identity();
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    G1Xyzz o = shfl_down_xyzz(y, d);
    if (lane < (uint32_t)d) xyzz_add(y, o);
  }
  t = y;
}
__device__ __forceinline__ G1Xyzz warp_sum(G1Xyzz y, uint32_t lane) {
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    G1Xyzz o = shfl_down_xyzz(y, d);
    if (lane < (uint32_t)d) xyzz_add(y, o);
  }
  return y;
}

// level 1: group g of MSM m = buckets [32g, 32g+32) -> s1[g] = sum X_l, t1[g] = sum l * X_l.
// Four lanes per group, eight consecutive buckets per lane with the running-sum trick (15 serial additions and no
// idle lanes), then two shuffle steps merge the four (s, t) pairs: t = t_a + t_b + width * s_b.
// (The one-bucket-per-lane version ran 10 additions on all 32 lanes for 31 useful ones and was throughput-bound at
// k = 17: 232 us per batch of 8 MSMs, profiles/r01_launches_proof_large_a.csv.)
__device__ __forceinline__ G1Xyzz xyzz_mul_pow2(G1Xyzz p, int log2w) {
#pragma unroll 1
  for (int i = 0; i < log2w; i++) p = xyzz_double(p);
  return p;
}
__global__ void __launch_bounds__(128) msm_bucket_l1_kernel(const G1Xyzz* __restrict__ buckets, uint32_t NB,
                                                            uint32_t n1, G1Xyzz* __restrict__ s1,
                                                            G1Xyzz* __restrict__ t1) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;   // eighth-of-a-warp index: buckets [8q, 8q+8)
  const uint32_t m = blockIdx.y;
  const uint32_t g = q >> 2, sub = q & 3;
  const bool live = g < n1;                                    // whole warps stay in the shuffles
  const G1Xyzz* B = buckets + (size_t)m * NB + (size_t)q * 8;
  G1Xyzz run = xyzz_identity(), acc = xyzz_identity();
  if (live) {
#pragma unroll 1
    for (int l = 7; l >= 1; l--) {
      if ((size_t)q * 8 + l < NB) {
        G1Xyzz x = B[l];
        xyzz_add(run, x);
      }
      xyzz_add(acc, run);                                      // acc = sum_{l >= 1} l * X_l
    }
    if ((size_t)q * 8 < NB) {
      G1Xyzz x = B[0];
      xyzz_add(run, x);                                        // run = sum X_l
    }
  }
  // merge pairs: (sub 0, 1) and (2, 3) with width 8, then (0, 2) with width 16
#pragma unroll 1
  for (int step = 0; step < 2; step++) {
    const int d = 1 << step;
    G1Xyzz so = shfl_down_xyzz(run, d);
    G1Xyzz to = shfl_down_xyzz(acc, d);
    if ((sub & (2 * d - 1)) == 0) {
      xyzz_add(run, so);
      xyzz_add(acc, to);
      so = xyzz_mul_pow2(so, 3 + step);
      xyzz_add(acc, so);
    }
  }
  if (live && sub == 0) {
    s1[(size_t)m * n1 + g] = run;
    t1[(size_t)m * n1 + g] = acc;
  }
}

// level 1, latency-optimal form (one bucket per lane, 10 serial additions, 32 lanes busy for 31 useful sums): used
// when the batch has few buckets and the kernel is bound by the addition chain, not by throughput
__global__ void __launch_bounds__(128) msm_bucket_l1_warp_kernel(const G1Xyzz* __restrict__ buckets, uint32_t NB,
                                                            uint32_t n1, G1Xyzz* __restrict__ s1,
                                                            G1Xyzz* __restrict__ t1) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t m = blockIdx.y;
  if (g >= n1) return;
  uint32_t b = g * 32 + lane;
  G1Xyzz x = (b < NB) ? buckets[(size_t)m * NB + b] : xyzz_identity();
  G1Xyzz s, t;
  warp_weighted(x, lane, s, t);
  if (lane == 0) {
    s1[(size_t)m * n1 + g] = s;
    t1[(size_t)m * n1 + g] = t;
  }
}

__device__ __forceinline__ G1Xyzz xyzz_mul32(G1Xyzz p) {
  for (int i = 0; i < 5; i++) p = xyzz_double(p);
  return p;
}

// level 2: CTA w of MSM m folds s1[32w..32w+32) -> (S2, T2) (warp 0) and sums t1[32w..32w+32) -> U (warp 1); the two
// chains (10 and 5 dependent additions) run side by side instead of back to back
__global__ void __launch_bounds__(64) msm_bucket_l2_kernel(const G1Xyzz* __restrict__ s1,
                                                           const G1Xyzz* __restrict__ t1, uint32_t n1,
                                                           G1Xyzz* __restrict__ l2out) {
  const uint32_t w = blockIdx.x, m = blockIdx.y, lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const uint32_t nw = gridDim.x;
  uint32_t i = w * 32 + lane;
  G1Xyzz* o = l2out + (size_t)m * 3 * nw;
  if (role == 0) {
    G1Xyzz x = (i < n1) ? s1[(size_t)m * n1 + i] : xyzz_identity();
    G1Xyzz s, t;
    warp_weighted(x, lane, s, t);
    if (lane == 0) {
      o[w] = s;
      o[nw + w] = t;
    }
  } else {
    G1Xyzz tt = (i < n1) ? t1[(size_t)m * n1 + i] : xyzz_identity();
    G1Xyzz u = warp_sum(tt, lane);
    if (lane == 0) o[2 * nw + w] = u;
  }
}

// finish: one 3-warp CTA per MSM folds the nw <= 32 (S2, T2, U) triples:
//   W(X) = sum U + 32 * ( sum T2 + 32 * t3 ),  result = W + S   (bucket `key` has weight key+1)
__global__ void __launch_bounds__(96) msm_finish_kernel(const G1Xyzz* __restrict__ l2out, uint32_t nw,
                                                        G1Jac* __restrict__ out) {
  __shared__ G1Xyzz fin[4];
  const uint32_t m = blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const G1Xyzz* o = l2out + (size_t)m * 3 * nw;
  if (wid == 0) {
    G1Xyzz x = (lane < nw) ? o[lane] : xyzz_identity();
    G1Xyzz s, t;
    warp_weighted(x, lane, s, t);
    if (lane == 0) {
      fin[0] = s;  // S: sum of all buckets
      fin[1] = t;  // t3
    }
  } else {
    G1Xyzz x = (lane < nw) ? o[wid * nw + lane] : xyzz_identity();
    x = warp_sum(x, lane);
    if (lane == 0) fin[1 + wid] = x;  // fin[2] = sum T2, fin[3] = sum U
  }
  __syncthreads();
  if (tid == 0) {
    G1Xyzz r = xyzz_mul32(fin[1]);
    xyzz_add(r, fin[2]);
    r = xyzz_mul32(r);
    xyzz_add(r, fin[3]);
    xyzz_add(r, fin[0]);
    out[m] = xyzz_to_jacobian(r);
  }
}

// ---- launchers called from msm_run (msm.cu) -------------------------------------------------------------
void msm_tail_serial_level(const uint32_t* keys, const G1Xyzz* pts, uint32_t slots, G1Xyzz* buckets, uint32_t* pkeys_out,
                           G1Xyzz* ppts_out, uint32_t T1, cudaStream_t st) {
  msm_serial_reduce_kernel<<<(T1 + 127) / 128, 128, 0, st>>>(keys, pts, slots, MSM_LEVEL1_K, buckets, pkeys_out, ppts_out, T1);
}
void msm_tail_warp_level(const uint32_t* keys, const G1Xyzz* pts, uint32_t slots, G1Xyzz* buckets, uint32_t* pkeys_out,
                         G1Xyzz* ppts_out, uint32_t nwarps, int fin, cudaStream_t st) {
  msm_warp_reduce_kernel<<<(nwarps + 3) / 4, 128, 0, st>>>(keys, pts, slots, buckets, pkeys_out, ppts_out, nwarps, fin);
}
void msm_tail_buckets(const G1Xyzz* buckets, uint32_t NB, uint32_t M, G1Xyzz* s1, G1Xyzz* t1, G1Xyzz* l2out, G1Jac* out,
                      cudaStream_t st) {
  uint32_t n1 = (NB + 31) / 32;
  // <= 2 warps per SM sub-partition in the one-bucket-per-lane form: the chain of 10 additions decides; beyond that
  // the 8-buckets-per-lane form (4x fewer lane-additions) wins
  if ((uint64_t)n1 * M <= 2 * 592)
    msm_bucket_l1_warp_kernel<<<dim3((n1 + 3) / 4, M), 128, 0, st>>>(buckets, NB, n1, s1, t1);
  else
    msm_bucket_l1_kernel<<<dim3((4 * n1 + 127) / 128, M), 128, 0, st>>>(buckets, NB, n1, s1, t1);
  uint32_t nw = (n1 + 31) / 32;
  msm_bucket_l2_kernel<<<dim3(nw, M), 64, 0, st>>>(s1, t1, n1, l2out);
  msm_finish_kernel<<<M, 96, 0, st>>>(l2out, nw, out);
}

}  // namespace zg
