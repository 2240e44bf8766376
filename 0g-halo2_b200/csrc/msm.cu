// Fixed-base signed-digit Pippenger MSM over BN254 G1 for sm_100a.
//
// Replaces halo2_proofs (tag v2023_04_20, un-vendored) `arithmetic::best_multiexp` as used by
// `ParamsKZG::commit` / `commit_lagrange` and the GWC opening proofs; call site
// /root/reference/src/wnn.rs:242-259 (create_proof) and :226-228 (keygen).
//
// B200-first design (not the upstream per-thread-chunk serial Pippenger):
//   * the SRS is fixed, so at load time we spend HBM (180 GB) on a window table
//     pts[w][i] = 2^(c*w) * G_i.  All W windows of a scalar then fall into ONE bucket set, which
//     removes the per-window bucket reductions and the 254 doublings of the window Horner.
//   * signed c-bit digits -> 2^(c-1) buckets per MSM; zero digits are skipped, so sparse /
//     small-valued advice columns cost proportionally less.
//   * (bucket, point) pairs are counting-sorted; accumulation is a load-balanced SEGMENTED
//     reduction over the sorted list (fixed-size chunks per thread), so a bucket that receives
//     20 000 points (value "1" in an advice column) costs the same per point as a uniform one.
//     Chunk-boundary partial sums are folded by a second serial level and then by warp-shuffle
//     segmented reductions.
//   * the weighted bucket sum  sum_b b * B_b  is a warp-shuffle suffix scan (32 buckets per
//     warp) followed by a one-CTA finish kernel.
//   * several MSMs over the same basis (one Fiat-Shamir round's commitments) share every launch.
#include <cstdlib>
#include "msm.cuh"

namespace zg {

__device__ __forceinline__ void ld_fq2(const G1Affine* p, Fq& x, Fq& y) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
  x.v[0] = a.x; x.v[1] = a.y; x.v[2] = a.z; x.v[3] = a.w;
  x.v[4] = b.x; x.v[5] = b.y; x.v[6] = b.z; x.v[7] = b.w;
  y.v[0] = c.x; y.v[1] = c.y; y.v[2] = c.z; y.v[3] = c.w;
  y.v[4] = d.x; y.v[5] = d.y; y.v[6] = d.z; y.v[7] = d.w;
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

// ---- window-table precompute (one-time, at SRS load) -----------------------------------
__global__ void msm_precompute_kernel(const G1Affine* __restrict__ base, G1Affine* __restrict__ table,
                                      uint32_t n, uint32_t c, uint32_t W) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = base[i];
  G1Xyzz q = xyzz_from_affine(p);
  for (uint32_t w = 0; w < W; w++) {
    G1Affine a;
    if (xyzz_is_identity(q)) {
      a.x = fp_zero<FqParams>();
      a.y = fp_zero<FqParams>();
    } else {
      // x = X/ZZ, y = Y/ZZZ with one inversion: i3 = 1/ZZZ, 1/ZZ = (ZZ*i3)^2  (ZZ^3 = ZZZ^2)
      Fq i3 = fp_inv(q.zzz);
      Fq t = fp_mul(q.zz, i3);
      a.x = fp_mul(q.x, fp_sqr(t));
      a.y = fp_mul(q.y, i3);
    }
    table[(size_t)w * n + i] = a;
    if (w + 1 < W) {
      // restart from the affine form: keeps zz = zzz = 1 so the c doublings stay cheap
      q = xyzz_from_affine(a);
      for (uint32_t d = 0; d < c; d++) q = xyzz_double(q);
    }
  }
}

// ---- signed digits --------------------------------------------------------------------
// canonical scalar limbs -> digit of window w given incoming carry; |digit| <= 2^(c-1)
__device__ __forceinline__ int32_t msm_digit(const uint32_t (&s)[8], uint32_t w, uint32_t c,
                                             uint32_t& carry) {
  uint32_t p = w * c;
  uint32_t limb = p >> 5, off = p & 31;
  uint32_t v = 0;
  if (limb < 8) {
    v = s[limb] >> off;
    if (off && limb + 1 < 8) v |= s[limb + 1] << (32 - off);
  }
  v &= (1u << c) - 1;
  v += carry;
  if (v > (1u << (c - 1))) {
    carry = 1;
    return (int32_t)v - (int32_t)(1u << c);
  }
  carry = 0;
  return (int32_t)v;
}

template <bool SCATTER>
__global__ void msm_digits_kernel(const Fr* __restrict__ scalars, size_t stride, uint32_t n_used,
                                  uint32_t n_tab, uint32_t c, uint32_t W, uint32_t NB,
                                  uint32_t* __restrict__ counters, uint32_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t m = blockIdx.y;
  if (i >= n_used) return;
  Fr sc = scalars[(size_t)m * stride + i];
  if (fp_is_zero(sc)) return;
  Fr can = fp_from_mont(sc);
  uint32_t carry = 0;
  for (uint32_t w = 0; w < W; w++) {
    int32_t d = msm_digit(can.v, w, c, carry);
    if (d == 0) continue;
    uint32_t neg = d < 0;
    uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
    uint32_t key = m * NB + (mag - 1);
    if (SCATTER) {
      uint32_t pos = atomicAdd(&counters[key], 1u);
      keys[pos] = key;
      vals[pos] = (w * n_tab + i) | (neg << 31);
    } else {
      atomicAdd(&counters[key], 1u);
    }
  }
}

// single-CTA exclusive scan of `cnt` counters: offsets[0..cnt] and a cursor copy
__global__ void __launch_bounds__(1024) msm_scan_kernel(const uint32_t* __restrict__ hist, uint32_t cnt,
                                                        uint32_t* __restrict__ offsets,
                                                        uint32_t* __restrict__ cursor) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t per = (cnt + 1023) / 1024;
  const uint32_t b = tid * per, e = min(b + per, cnt);
  uint32_t sum = 0;
  for (uint32_t i = b; i < e; i++) sum += hist[i];
  // block exclusive scan of the per-thread sums
  uint32_t incl = sum;
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane >= d) incl += o;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = warp_sums[lane];
    uint32_t wi = ws;
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
      if ((int)lane >= d) wi += o;
    }
    warp_sums[lane] = wi - ws;
    if (lane == 31) carry_s = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[wid] + incl - sum;
  for (uint32_t i = b; i < e; i++) {
    offsets[i] = run;
    cursor[i] = run;
    run += hist[i];
  }
  if (tid == 0) offsets[cnt] = carry_s;
}

// ---- serial segmented reduction over the sorted list -----------------------------------
// LEVEL0: entries are (key, table index|sign) and are gathered from the affine window table.
// !LEVEL0: entries are (key, XYZZ partial sum).
// Thread t owns entries [t*K, t*K+K).  Runs strictly inside the chunk go straight to their bucket
// (nobody else holds that key); the first and last runs go to partial slots 2t, 2t+1.
template <bool LEVEL0>
__global__ void __launch_bounds__(128) msm_serial_reduce_kernel(
    const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
    const G1Xyzz* __restrict__ pts_in, const uint32_t* __restrict__ count_ptr, uint32_t count_static,
    const G1Affine* __restrict__ table, uint32_t K, G1Xyzz* __restrict__ buckets,
    uint32_t* __restrict__ pkeys, G1Xyzz* __restrict__ ppts, uint32_t nthreads) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  const uint32_t L = count_ptr ? *count_ptr : count_static;
  const uint64_t start64 = (uint64_t)t * K;
  if (start64 >= L) {
    pkeys[2 * t] = MSM_INVALID_KEY;
    pkeys[2 * t + 1] = MSM_INVALID_KEY;
    return;
  }
  const uint32_t start = (uint32_t)start64;
  const uint32_t end = (start64 + K < L) ? start + K : L;
  uint32_t cur = keys[start];
  uint32_t e = start;
  if (!LEVEL0) {
    // the list may end in invalid slots
    if (cur == MSM_INVALID_KEY) {
      pkeys[2 * t] = MSM_INVALID_KEY;
      pkeys[2 * t + 1] = MSM_INVALID_KEY;
      return;
    }
  }
  G1Xyzz acc = xyzz_identity();
  uint32_t nruns = 0;
  for (; e < end; e++) {
    uint32_t k = keys[e];
    if (!LEVEL0 && k == MSM_INVALID_KEY) break;
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
    if (LEVEL0) {
      uint32_t v = vals[e];
      Fq x, y;
      ld_fq2(table + (v & 0x7fffffffu), x, y);
      if (fp_is_zero(x) && fp_is_zero(y)) continue;
      if (v >> 31) y = fp_neg(y);
      xyzz_madd(acc, x, y);
    } else {
      G1Xyzz p = pts_in[e];
      xyzz_add(acc, p);
    }
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

// level 1: warp g of MSM m reduces buckets [32g, 32g+32) to (s1, t1)
__global__ void __launch_bounds__(128) msm_bucket_l1_kernel(const G1Xyzz* __restrict__ buckets, uint32_t NB,
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

// level 2: warp w of MSM m folds s1[32w..32w+32) -> (S2, T2) and sums t1[32w..32w+32) -> U
__global__ void __launch_bounds__(32) msm_bucket_l2_kernel(const G1Xyzz* __restrict__ s1,
                                                           const G1Xyzz* __restrict__ t1, uint32_t n1,
                                                           G1Xyzz* __restrict__ l2out) {
  const uint32_t w = blockIdx.x, m = blockIdx.y, lane = threadIdx.x;
  const uint32_t nw = gridDim.x;
  uint32_t i = w * 32 + lane;
  G1Xyzz x = (i < n1) ? s1[(size_t)m * n1 + i] : xyzz_identity();
  G1Xyzz tt = (i < n1) ? t1[(size_t)m * n1 + i] : xyzz_identity();
  G1Xyzz s, t;
  warp_weighted(x, lane, s, t);
  G1Xyzz u = warp_sum(tt, lane);
  if (lane == 0) {
    G1Xyzz* o = l2out + (size_t)m * 3 * nw;
    o[w] = s;
    o[nw + w] = t;
    o[2 * nw + w] = u;
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

// ---- host side ---------------------------------------------------------------------------
uint32_t msm_pick_c(uint32_t k) {
  if (const char* e = getenv("ZG_MSM_C")) {
    int v = atoi(e);
    if (v >= 6 && v <= 16) return (uint32_t)v;
  }
  int c = (int)k - 2;
  if (c < 8) c = 8;
  if (c > 16) c = 16;
  return (uint32_t)c;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

MsmWorkspaceLayout msm_workspace_layout(uint32_t n, uint32_t c, uint32_t W, uint32_t M) {
  MsmWorkspaceLayout l;
  l.M = M;
  l.NB = 1u << (c - 1);
  uint64_t lmax = (uint64_t)n * W * M;
  l.L_max = (uint32_t)lmax;
  l.K0 = lmax >= (2u << 20) ? 32 : 16;
  l.T0 = (uint32_t)((lmax + l.K0 - 1) / l.K0);
  l.slots_a = 2 * l.T0;
  // level 1 (serial, K = 16) or first warp level consumes slots_a
  uint32_t t1s = (l.slots_a + 15) / 16, t1w = (l.slots_a + 31) / 32;
  l.slots_b = 2 * (t1s > t1w ? t1s : t1w);
  size_t o = 0;
  size_t cnt = (size_t)M * l.NB;
  l.off_hist = o; o = align_up(o + (cnt + 1) * 4);
  l.off_cursor = o; o = align_up(o + (cnt + 1) * 4);
  l.off_offsets = o; o = align_up(o + (cnt + 1) * 4);
  l.off_keys = o; o = align_up(o + (size_t)l.L_max * 4);
  l.off_vals = o; o = align_up(o + (size_t)l.L_max * 4);
  l.off_buckets = o; o = align_up(o + cnt * sizeof(G1Xyzz));
  l.off_pkeys_a = o; o = align_up(o + (size_t)l.slots_a * 4);
  l.off_ppts_a = o; o = align_up(o + (size_t)l.slots_a * sizeof(G1Xyzz));
  l.off_pkeys_b = o; o = align_up(o + (size_t)l.slots_b * 4);
  l.off_ppts_b = o; o = align_up(o + (size_t)l.slots_b * sizeof(G1Xyzz));
  size_t n1 = (l.NB + 31) / 32;
  l.off_s1 = o; o = align_up(o + (size_t)M * n1 * sizeof(G1Xyzz));
  l.off_t1 = o; o = align_up(o + (size_t)M * n1 * sizeof(G1Xyzz));
  l.off_l2 = o; o = align_up(o + (size_t)M * 3 * ((n1 + 31) / 32) * sizeof(G1Xyzz));
  l.bytes = o;
  return l;
}

cudaError_t msm_precompute_table(const G1Affine* base, uint32_t n, uint32_t c, uint32_t W,
                                 G1Affine* table, cudaStream_t stream) {
  msm_precompute_kernel<<<(n + 63) / 64, 64, 0, stream>>>(base, table, n, c, W);
  return cudaGetLastError();
}

cudaError_t msm_run(const MsmTable& tb, const Fr* scalars, size_t stride, uint32_t n_used, uint32_t M,
                    G1Jac* out, uint8_t* ws, const MsmWorkspaceLayout& l, cudaStream_t st, uint64_t* nl) {
  uint64_t launches = 0;
  const uint32_t NB = l.NB, cnt = M * NB;
  uint32_t* hist = (uint32_t*)(ws + l.off_hist);
  uint32_t* cursor = (uint32_t*)(ws + l.off_cursor);
  uint32_t* offsets = (uint32_t*)(ws + l.off_offsets);
  uint32_t* keys = (uint32_t*)(ws + l.off_keys);
  uint32_t* vals = (uint32_t*)(ws + l.off_vals);
  G1Xyzz* buckets = (G1Xyzz*)(ws + l.off_buckets);
  uint32_t* pk[2] = {(uint32_t*)(ws + l.off_pkeys_a), (uint32_t*)(ws + l.off_pkeys_b)};
  G1Xyzz* pp[2] = {(G1Xyzz*)(ws + l.off_ppts_a), (G1Xyzz*)(ws + l.off_ppts_b)};
  G1Xyzz* s1 = (G1Xyzz*)(ws + l.off_s1);
  G1Xyzz* t1 = (G1Xyzz*)(ws + l.off_t1);

  cudaMemsetAsync(hist, 0, (size_t)(cnt + 1) * 4, st);
  cudaMemsetAsync(buckets, 0, (size_t)cnt * sizeof(G1Xyzz), st);
  dim3 dgrid((n_used + 127) / 128, M);
  launches++;
  msm_digits_kernel<false><<<dgrid, 128, 0, st>>>(scalars, stride, n_used, tb.n, tb.c, tb.W, NB, hist,
                                                  nullptr, nullptr);
  launches++;
  msm_scan_kernel<<<1, 1024, 0, st>>>(hist, cnt, offsets, cursor);
  launches++;
  msm_digits_kernel<true><<<dgrid, 128, 0, st>>>(scalars, stride, n_used, tb.n, tb.c, tb.W, NB, cursor,
                                                 keys, vals);
  // level 0: serial chunks over the sorted list (length offsets[cnt], read on device)
  uint32_t T0 = l.T0;
  launches++;
  msm_serial_reduce_kernel<true><<<(T0 + 127) / 128, 128, 0, st>>>(
      keys, vals, nullptr, offsets + cnt, 0, tb.pts, l.K0, buckets, pk[0], pp[0], T0);
  uint32_t slots = 2 * T0;
  int cur = 0;
  if (slots > 8192) {
    uint32_t T1 = (slots + 15) / 16;
    launches++;
    msm_serial_reduce_kernel<false><<<(T1 + 127) / 128, 128, 0, st>>>(
        pk[0], nullptr, pp[0], nullptr, slots, nullptr, 16, buckets, pk[1], pp[1], T1);
    slots = 2 * T1;
    cur = 1;
  }
  for (;;) {
    uint32_t nwarps = (slots + 31) / 32;
    int fin = nwarps == 1;
    launches++;
    msm_warp_reduce_kernel<<<(nwarps + 3) / 4, 128, 0, st>>>(pk[cur], pp[cur], slots, buckets,
                                                            pk[cur ^ 1], pp[cur ^ 1], nwarps, fin);
    if (fin) break;
    slots = 2 * nwarps;
    cur ^= 1;
  }
  uint32_t n1 = (NB + 31) / 32;
  dim3 g1((n1 + 3) / 4, M);
  launches++;
  msm_bucket_l1_kernel<<<g1, 128, 0, st>>>(buckets, NB, n1, s1, t1);
  uint32_t nw = (n1 + 31) / 32;
  G1Xyzz* l2out = (G1Xyzz*)(ws + l.off_l2);
  launches++;
  msm_bucket_l2_kernel<<<dim3(nw, M), 32, 0, st>>>(s1, t1, n1, l2out);
  launches++;
  msm_finish_kernel<<<M, 96, 0, st>>>(l2out, nw, out);
  if (nl) *nl += launches;
  return cudaGetLastError();
}

}  // namespace zg
