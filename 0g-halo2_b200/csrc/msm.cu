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
//   * (bucket, point) pairs are counting-sorted by CTAs that keep the bucket counters in shared memory (no global
//     atomics; msm_digits_kernel below); accumulation is a load-balanced SEGMENTED reduction over the sorted list
//     (fixed-size chunks per thread), so a bucket that receives 20 000 points (value "1" in an advice column) costs
//     the same per point as a uniform one.  Chunk-boundary partial sums are folded by a second serial level and then
//     by warp-shuffle segmented reductions (msm_tail.cu).
//   * the weighted bucket sum  sum_b (b + 1) * B_b  has a latency-optimal and a throughput-optimal first level
//     (msm_tail.cu), a warp-parallel second level and a one-CTA finish kernel.
//   * several MSMs (one Fiat-Shamir round's commitments) share every launch, and a batch may mix the two bases.
// The level-0 accumulation is 100 KB of SASS with the multiplier inlined and stalls on instruction fetch
// (ncu: stalled_no_instruction 3.0 per issue, profiles/r01_ncu_full_msm_accumulate_small_proof.txt):
// outline it (field.cuh).
#ifndef ZG_MSM_INLINE_MUL
#define ZG_FP_MUL_NOINLINE 1
#endif
#include <atomic>
#include <cstdlib>
#include "msm.cuh"
#include "scan.cuh"

#ifndef ZG_MSM_MADD_DEFAULT
#define ZG_MSM_MADD_DEFAULT 1
#endif

namespace zg {

__device__ __forceinline__ void ld_fq2(const G1Affine* p, Fq& x, Fq& y) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
  x.v[0] = a.x; x.v[1] = a.y; x.v[2] = a.z; x.v[3] = a.w;
  x.v[4] = b.x; x.v[5] = b.y; x.v[6] = b.z; x.v[7] = b.w;
  y.v[0] = c.x; y.v[1] = c.y; y.v[2] = c.z; y.v[3] = c.w;
  y.v[4] = d.x; y.v[5] = d.y; y.v[6] = d.z; y.v[7] = d.w;
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

// Counting sort of the digits by bucket WITHOUT global atomics: CTA (j, m) owns scalars [j*chunk, (j+1)*chunk) of
// MSM m and keeps that MSM's 2^(c-1) bucket counters in shared memory.
//   count pass   : shared-memory histogram of the chunk, flushed to H[m][j][bucket]
//   offsets      : msm_hist_total / scan / msm_hist_offsets turn H into the first output slot of (m, bucket, j)
//   scatter pass : the CTA reloads its slice of H as cursors and writes packed (bucket key, table index | sign)
// The global-atomic version spent 3.6 ms of a 30 ms k = 17 proof here (profiles/r01_launches_proof_large_a.csv).
constexpr int DG_THREADS = 512;

template <bool SCATTER>
__global__ void __launch_bounds__(DG_THREADS) msm_digits_kernel(const Fr* __restrict__ scalars, size_t stride, uint32_t n_used,
                                                                uint32_t n_tab, uint32_t c, uint32_t W, uint32_t NB, uint32_t chunk,
                                                                uint32_t* __restrict__ H, uint2* __restrict__ entries) {
  extern __shared__ uint32_t dg_sh[];   // NB counters / cursors
  const uint32_t m = blockIdx.y, j = blockIdx.x, J = gridDim.x;
  uint32_t* Hb = H + ((size_t)m * J + j) * NB;
  for (uint32_t b = threadIdx.x; b < NB; b += DG_THREADS) dg_sh[b] = SCATTER ? Hb[b] : 0u;
  __syncthreads();
  const uint32_t lo = j * chunk;
  const uint32_t hi = min(lo + chunk, n_used);
  for (uint32_t i = lo + threadIdx.x; i < hi; i += DG_THREADS) {
    Fr sc = scalars[(size_t)m * stride + i];
    if (fp_is_zero(sc)) continue;
    Fr can = fp_from_mont(sc);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < W; w++) {
      int32_t d = msm_digit(can.v, w, c, carry);
      if (d == 0) continue;
      uint32_t neg = d < 0;
      uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
      uint32_t pos = atomicAdd(&dg_sh[mag - 1], 1u);
      if (SCATTER) entries[pos] = make_uint2(m * NB + (mag - 1), (w * n_tab + i) | (neg << 31));
    }
  }
  if (!SCATTER) {
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < NB; b += DG_THREADS) Hb[b] = dg_sh[b];
  }
}

// total[m*NB + b] = sum_j H[m][j][b]
__global__ void msm_hist_total_kernel(const uint32_t* __restrict__ H, uint32_t J, uint32_t NB, uint32_t cnt,
                                      uint32_t* __restrict__ total) {
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cnt) return;
  const uint32_t m = idx / NB, b = idx - m * NB;
  const uint32_t* p = H + (size_t)m * J * NB + b;
  uint32_t s = 0;
#pragma unroll 8
  for (uint32_t j = 0; j < J; j++) s += __ldg(p + (size_t)j * NB);
  total[idx] = s;
}
// H[m][j][b] <- offsets[m*NB + b] + sum_{j' < j} H[m][j'][b]
// (loads are issued eight at a time: the in-place update would otherwise serialise on one load latency per chunk)
__global__ void msm_hist_offsets_kernel(uint32_t* __restrict__ H, uint32_t J, uint32_t NB, uint32_t cnt,
                                        const uint32_t* __restrict__ offsets) {
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cnt) return;
  const uint32_t m = idx / NB, b = idx - m * NB;
  uint32_t* p = H + (size_t)m * J * NB + b;
  uint32_t run = offsets[idx];
  for (uint32_t j0 = 0; j0 < J; j0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = (j0 + u < J) ? p[(size_t)(j0 + u) * NB] : 0u;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (j0 + u < J) p[(size_t)(j0 + u) * NB] = run;
      run += v[u];
    }
  }
}

// ---- serial segmented reduction over the sorted list -----------------------------------
// Level 0 of the accumulation: entries are packed (bucket key, table index | sign) pairs, gathered from the
// affine window table.  (The XYZZ-partial levels live in msm_tail.cu.)
// Thread t owns entries [t*K, t*K+K).  Runs strictly inside the chunk go straight to their bucket
// (nobody else holds that key); the first and last runs go to partial slots 2t, 2t+1.
template <bool LAZY>
__global__ void __launch_bounds__(128, 4) msm_accumulate_kernel(
    const uint2* __restrict__ entries, const uint32_t* __restrict__ count_ptr, const G1Affine* __restrict__ table,
    const G1Affine* __restrict__ alt_table, uint32_t alt_mask, uint32_t log_nb,
    uint32_t K, G1Xyzz* __restrict__ buckets, uint32_t* __restrict__ pkeys, G1Xyzz* __restrict__ ppts, uint32_t nthreads) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  const uint32_t L = *count_ptr;
  // equal shares of the REAL list (its length is only known on the device: zero digits were skipped), never below the
  // minimum chunk K: the grid is sized to whole waves of resident CTAs, so every SM finishes at the same time
  const uint32_t share = (uint32_t)(((uint64_t)L + nthreads - 1) / nthreads);
  if (share > K) K = share;
  const uint64_t start64 = (uint64_t)t * K;
  if (start64 >= L) {
    pkeys[2 * t] = MSM_INVALID_KEY;
    pkeys[2 * t + 1] = MSM_INVALID_KEY;
    return;
  }
  const uint32_t start = (uint32_t)start64;
  const uint32_t end = (start64 + K < L) ? start + K : L;
  uint2 ent = __ldg(entries + start);
  uint32_t cur = ent.x;
  G1Xyzz acc = xyzz_identity();
  uint32_t nruns = 0;
  for (uint32_t e = start; e < end; e++) {
    const uint2 nxt = (e + 1 < end) ? __ldg(entries + e + 1) : ent;   // prefetch the next pair
    if (ent.x != cur) {
      if (nruns == 0) {
        pkeys[2 * t] = cur;
        ppts[2 * t] = acc;
      } else {
        buckets[cur] = acc;
      }
      nruns++;
      cur = ent.x;
      acc = xyzz_identity();
    }
    Fq x, y;
    const G1Affine* tb = ((alt_mask >> (ent.x >> log_nb)) & 1u) ? alt_table : table;
    ld_fq2(tb + (ent.y & 0x7fffffffu), x, y);
    if (!(fp_is_zero(x) && fp_is_zero(y))) {
      if (ent.y >> 31) y = fp_neg(y);
      xyzz_madd_t<LAZY>(acc, x, y);
    }
    ent = nxt;
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

void msm_tail_serial_level(const uint32_t* keys, const G1Xyzz* pts, uint32_t slots, G1Xyzz* buckets, uint32_t* pkeys_out,
                           G1Xyzz* ppts_out, uint32_t T1, cudaStream_t st);
void msm_tail_warp_level(const uint32_t* keys, const G1Xyzz* pts, uint32_t slots, G1Xyzz* buckets, uint32_t* pkeys_out,
                         G1Xyzz* ppts_out, uint32_t nwarps, int fin, cudaStream_t st);
void msm_tail_buckets(const G1Xyzz* buckets, uint32_t NB, uint32_t M, G1Xyzz* s1, G1Xyzz* t1, G1Xyzz* l2out, G1Jac* out,
                      cudaStream_t st);

// ---- host side ---------------------------------------------------------------------------
// threads of one full wave of msm_accumulate_kernel on the current device (resident CTAs per SM x SMs x 128)
static uint32_t msm_accumulate_wave_threads() {
  static std::atomic<uint32_t> cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  uint32_t v = cached[dev & 63].load(std::memory_order_acquire);
  if (v) return v;
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msm_accumulate_kernel<true>, 128, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (per_sm < 1) per_sm = 3;
  if (sms < 1) sms = 148;
  v = (uint32_t)per_sm * (uint32_t)sms * 128u;
  cached[dev & 63].store(v, std::memory_order_release);
  return v;
}

// mixed addition of the level-0 accumulation: 1 (default) = dedicated squarings + one lazily reduced two-product for Y3
// (curve.cuh::xyzz_madd_t<true>), 0 = ten fully reduced products.  Same bucket contents bit for bit; ZG_MSM_MADD selects.
static bool msm_madd_lazy() {
  static const int v = [] {
    const char* e = getenv("ZG_MSM_MADD");
    return e ? atoi(e) : ZG_MSM_MADD_DEFAULT;
  }();
  return v != 0;
}

uint32_t msm_pick_c(uint32_t k) {
  if (const char* e = getenv("ZG_MSM_C")) {
    int v = atoi(e);
    if (v >= 6 && v <= 16) return (uint32_t)v;
  }
  // k - 2 up to 2^16 points; 16 from 2^17 on (k = 17 measured 61.5 vs 61.0 proofs/s and 19.3 vs 19.6 ms against c = 15,
  // profiles/r02_ab1_*.json: one window less per scalar outweighs the doubled bucket reduction)
  int c = k >= 17 ? 16 : (int)k - 2;
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
  // accumulation geometry: K0 entries per thread (32 for long lists: fewer chunk-boundary partials for the tail; measured
  // best of 8/16/32/64 for the k = 15 and k = 17 proofs, profiles/r01_notes.md).  ZG_MSM_WAVES caps the grid at that many
  // full waves of resident CTAs instead, the kernel then giving every thread an equal share of the real list -- measured
  // slower (one wave: 0.75 vs 0.83 of the IMAD peak at 2^20: no back-fill when SMs finish unevenly), kept for tuning.
  l.K0 = lmax >= (2u << 20) ? 32 : 16;
  if (const char* e = getenv("ZG_MSM_K0")) {          // tuning override: minimum entries per thread
    int v = atoi(e);
    if (v >= 4 && v <= 256) l.K0 = (uint32_t)v;
  }
  uint64_t t_cap = ~0ull;
  if (const char* e = getenv("ZG_MSM_WAVES")) {
    int v = atoi(e);
    if (v >= 1 && v <= 64) t_cap = (uint64_t)msm_accumulate_wave_threads() * (uint64_t)v;
  }
  const uint64_t t_dense = (lmax + l.K0 - 1) / l.K0;
  l.T0 = (uint32_t)(t_dense < t_cap ? t_dense : t_cap);
  // Few-wave launches (the commitment rounds of a proof: 1.1 - 4.3 waves of resident CTAs at K0 entries per thread) lose
  // up to half of their last wave.  When the dense grid is w waves and ceil(w) / w exceeds what a grid of exactly
  // floor(w) waves loses to SMs finishing unevenly (measured with ZG_MSM_WAVES at 2^20: 10 % at one wave, 6 % at two,
  // 2.5 % at four), launch floor(w) whole waves instead: the kernel then hands every thread L / threads > K0 entries.
  // Long lists (w > 6) keep the dense grid -- its back-fill wins there.  ZG_MSM_AUTOWAVES=0 disables the rule.
  static const bool auto_waves = [] {
    const char* e = getenv("ZG_MSM_AUTOWAVES");
    return !e || atoi(e) != 0;
  }();
  if (auto_waves && t_cap == ~0ull) {
    const uint64_t wt = msm_accumulate_wave_threads();
    const uint64_t nw = t_dense / wt;
    if (nw >= 1 && nw <= 6 && t_dense % wt != 0) {
      static const double loss[7] = {0, 0.10, 0.06, 0.04, 0.03, 0.025, 0.02};
      const double w = (double)t_dense / (double)wt;
      if ((double)(nw + 1) / w > 1.0 + loss[nw]) l.T0 = (uint32_t)(nw * wt);
    }
  }
  l.slots_a = 2 * l.T0;
  // level 1 (serial, K = 16) or first warp level consumes slots_a
  uint32_t t1s = (l.slots_a + MSM_LEVEL1_K - 1) / MSM_LEVEL1_K, t1w = (l.slots_a + 31) / 32;
  l.slots_b = 2 * (t1s > t1w ? t1s : t1w);
  size_t o = 0;
  size_t cnt = (size_t)M * l.NB;
  // digit-sort geometry: J chunks per MSM, about four CTAs per SM over the whole batch
  uint32_t J = (592 + M - 1) / M;
  // a CTA zeroes and flushes all NB counters: keep its chunk at >= NB/8 scalars (>= 2 W digits per counter)
  uint32_t min_chunk = l.NB / 8 > 256 ? l.NB / 8 : 256;
  uint32_t jmax = (n + min_chunk - 1) / min_chunk;
  if (J > jmax) J = jmax;
  if (J < 1) J = 1;
  l.chunk = (n + J - 1) / J;
  l.J = (n + l.chunk - 1) / l.chunk;
  l.off_hist = o; o = align_up(o + (cnt + 1) * 4);                 // per-bucket totals
  l.off_cursor = o; o = align_up(o + (size_t)cnt * l.J * 4);       // H[m][j][bucket]
  l.off_offsets = o; o = align_up(o + (cnt + 1) * 4);
  l.off_keys = o; o = align_up(o + (size_t)l.L_max * 8);   // packed (key, val) entries
  l.off_vals = o;
  l.off_buckets = o; o = align_up(o + cnt * sizeof(G1Xyzz));
  l.off_pkeys_a = o; o = align_up(o + (size_t)l.slots_a * 4);
  l.off_ppts_a = o; o = align_up(o + (size_t)l.slots_a * sizeof(G1Xyzz));
  l.off_pkeys_b = o; o = align_up(o + (size_t)l.slots_b * 4);
  l.off_ppts_b = o; o = align_up(o + (size_t)l.slots_b * sizeof(G1Xyzz));
  size_t n1 = (l.NB + 31) / 32;
  l.off_s1 = o; o = align_up(o + (size_t)M * n1 * sizeof(G1Xyzz));
  l.off_t1 = o; o = align_up(o + (size_t)M * n1 * sizeof(G1Xyzz));
  l.off_l2 = o; o = align_up(o + (size_t)M * 3 * ((n1 + 31) / 32) * sizeof(G1Xyzz));
  l.off_scan = o; o = align_up(o + scan_scratch_words((uint32_t)cnt, 1) * 4);
  l.bytes = o;
  return l;
}

cudaError_t msm_precompute_table(const G1Affine* base, uint32_t n, uint32_t c, uint32_t W,
                                 G1Affine* table, cudaStream_t stream) {
  msm_precompute_kernel<<<(n + 63) / 64, 64, 0, stream>>>(base, table, n, c, W);
  return cudaGetLastError();
}

cudaError_t msm_run(const MsmTable& tb, const Fr* scalars, size_t stride, uint32_t n_used, uint32_t M,
                    G1Jac* out, uint8_t* ws, const MsmWorkspaceLayout& l, cudaStream_t st, uint64_t* nl, MsmProbe* probe,
                    const G1Affine* alt_pts, uint32_t alt_mask) {
  uint64_t launches = 0;
  {
    // up to 2^15 bucket counters (c = 16) in shared memory: opt in once per device
    static std::atomic<uint64_t> attr_devices{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_devices.load(std::memory_order_acquire) >> (dev & 63) & 1)) {
      cudaFuncSetAttribute(msm_digits_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
      cudaFuncSetAttribute(msm_digits_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
      attr_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  const uint32_t NB = l.NB, cnt = M * NB;
  uint32_t* hist = (uint32_t*)(ws + l.off_hist);
  uint32_t* cursor = (uint32_t*)(ws + l.off_cursor);
  uint32_t* offsets = (uint32_t*)(ws + l.off_offsets);
  uint2* entries = (uint2*)(ws + l.off_keys);
  G1Xyzz* buckets = (G1Xyzz*)(ws + l.off_buckets);
  uint32_t* pk[2] = {(uint32_t*)(ws + l.off_pkeys_a), (uint32_t*)(ws + l.off_pkeys_b)};
  G1Xyzz* pp[2] = {(G1Xyzz*)(ws + l.off_ppts_a), (G1Xyzz*)(ws + l.off_ppts_b)};
  G1Xyzz* s1 = (G1Xyzz*)(ws + l.off_s1);
  G1Xyzz* t1 = (G1Xyzz*)(ws + l.off_t1);

  cudaMemsetAsync(buckets, 0, (size_t)cnt * sizeof(G1Xyzz), st);
  // the layout was sized for n_used == l.n; fewer scalars simply leave trailing chunks empty
  uint32_t* H = cursor;
  dim3 dgrid(l.J, M);
  const size_t dsm = (size_t)NB * 4;
  launches++;
  msm_digits_kernel<false><<<dgrid, DG_THREADS, dsm, st>>>(scalars, stride, n_used, tb.n, tb.c, tb.W, NB, l.chunk, H, nullptr);
  launches++;
  msm_hist_total_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(H, l.J, NB, cnt, hist);
  {
    ScanJobs sj{};
    sj.in[0] = hist; sj.out[0] = offsets;
    scan_excl_u32(sj, 1, cnt, (uint32_t*)(ws + l.off_scan), st, LaunchCounter{&launches});
  }
  launches++;
  msm_hist_offsets_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(H, l.J, NB, cnt, offsets);
  launches++;
  msm_digits_kernel<true><<<dgrid, DG_THREADS, dsm, st>>>(scalars, stride, n_used, tb.n, tb.c, tb.W, NB, l.chunk, H, entries);
  // level 0: serial chunks over the sorted list (length offsets[cnt], read on device)
  uint32_t T0 = l.T0;
  launches++;
  const bool probing = probe && probe->on && probe->used < probe->cap;
  if (probing) {
    while (probe->ev.size() < 2 * (probe->used + 1)) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      probe->ev.push_back(e);
    }
    cudaEventRecord(probe->ev[2 * probe->used], st);
  }
  if (msm_madd_lazy())
    msm_accumulate_kernel<true><<<(T0 + 127) / 128, 128, 0, st>>>(entries, offsets + cnt, tb.pts, alt_pts ? alt_pts : tb.pts,
                                                                alt_pts ? alt_mask : 0u, tb.c - 1, l.K0, buckets, pk[0], pp[0], T0);
  else
    msm_accumulate_kernel<false><<<(T0 + 127) / 128, 128, 0, st>>>(entries, offsets + cnt, tb.pts, alt_pts ? alt_pts : tb.pts,
                                                                 alt_pts ? alt_mask : 0u, tb.c - 1, l.K0, buckets, pk[0], pp[0], T0);
  if (probing) {
    cudaEventRecord(probe->ev[2 * probe->used + 1], st);
    cudaMemcpyAsync(probe->counts + probe->used, offsets + cnt, 4, cudaMemcpyDeviceToHost, st);
    probe->used++;
  }
  uint32_t slots = 2 * T0;
  int cur = 0;
  if (slots > 8192) {
    uint32_t T1 = (slots + MSM_LEVEL1_K - 1) / MSM_LEVEL1_K;
    launches++;
    msm_tail_serial_level(pk[0], pp[0], slots, buckets, pk[1], pp[1], T1, st);
    slots = 2 * T1;
    cur = 1;
  }
  for (;;) {
    uint32_t nwarps = (slots + 31) / 32;
    int fin = nwarps == 1;
    launches++;
    msm_tail_warp_level(pk[cur], pp[cur], slots, buckets, pk[cur ^ 1], pp[cur ^ 1], nwarps, fin, st);
    if (fin) break;
    slots = 2 * nwarps;
    cur ^= 1;
  }
  G1Xyzz* l2out = (G1Xyzz*)(ws + l.off_l2);
  launches += 3;
  msm_tail_buckets(buckets, NB, M, s1, t1, l2out, out, st);
  if (nl) *nl += launches;
  return cudaGetLastError();
}

}  // namespace zg
