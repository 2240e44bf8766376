// Lookup argument: permute_expression_pair on the device.
//
// Replaces halo2_proofs v2023_04_20 (un-vendored; /root/reference/Cargo.toml:21-25)
// `plonk::lookup::prover::permute_expression_pair` -- upstream: single-threaded `sort()` of the
// compressed inputs, a BTreeMap multiset of the table, then "first occurrence takes its own value,
// remaining table values ascending go to the repeated rows from the highest row down"
// (SURVEY.md Appendix B.3).  Call site: create_proof, /root/reference/src/wnn.rs:242-259.
//
// B200-first restatement that yields the same two columns without sorting the inputs:
//   1. LSD radix sort of the (canonical) table values only (top 64 bits first, verified, full
//      256-bit fallback), unique values U with multiplicities cntT;
//   2. every input is ranked by binary search in U (an input that is not in the table raises the
//      constraint-system failure flag) and counted: cntA;
//   3. exclusive scans give the start of every value's run in A', the ascending list of left-over
//      table copies, and the index of every repeated row, from which each row of S' is a direct
//      look-up (no serial walk).
#include "lookup.cuh"

namespace zg {

namespace {

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
// canonical integers: compare from the most significant limb
__device__ __forceinline__ int cmp256(const Fr& a, const Fr& b) {
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] < b.v[i]) return -1;
    if (a.v[i] > b.v[i]) return 1;
  }
  return 0;
}

__global__ void k_canonical(const Fr* __restrict__ in, Fr* __restrict__ out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(out + i, fp_from_mont(ldf(in + i)));
}
__global__ void k_iota(uint32_t* a, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// ---- single-CTA exclusive scan of u32 (out has n+1 entries, out[n] = total) -----------------------
__global__ void __launch_bounds__(1024) k_scan_excl_u32(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t total_s;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t per = (n + 1023) / 1024;
  const uint32_t b = min(tid * per, n), e = min(b + per, n);
  uint32_t sum = 0;
  for (uint32_t i = b; i < e; i++) sum += in[i];
  uint32_t incl = sum;
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane >= d) incl += o;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = warp_sums[lane], wi = ws;
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
      if ((int)lane >= d) wi += o;
    }
    warp_sums[lane] = wi - ws;
    if (lane == 31) total_s = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[wid] + incl - sum;
  for (uint32_t i = b; i < e; i++) {
    uint32_t v = in[i];   // in and out may not alias
    out[i] = run;
    run += v;
  }
  if (tid == 0) out[n] = total_s;
}

// ---- LSD radix sort of indices by one key byte -----------------------------------------------------
constexpr int RS_CHUNK = 64;     // elements per thread
constexpr int RS_THREADS = 32;   // threads per CTA (one u32[256] offset table per thread in smem)

__device__ __forceinline__ uint32_t key_byte(const Fr* keys, uint32_t idx, uint32_t byte) {
  return (keys[idx].v[byte >> 2] >> ((byte & 3) * 8)) & 0xff;
}

// hist[d * T + t] = number of elements of chunk t whose digit is d
__global__ void __launch_bounds__(RS_THREADS) k_rs_count(const Fr* __restrict__ keys, const uint32_t* __restrict__ idx, uint32_t n,
                                                         uint32_t byte, uint32_t T, uint32_t* __restrict__ hist) {
  __shared__ uint32_t cnt[256][RS_THREADS];
  const uint32_t t = blockIdx.x * RS_THREADS + threadIdx.x;
  for (int d = 0; d < 256; d++) cnt[d][threadIdx.x] = 0;
  const uint32_t b = t * RS_CHUNK;
  if (t < T) {
    const uint32_t e = min(b + RS_CHUNK, n);
    for (uint32_t i = b; i < e; i++) cnt[key_byte(keys, idx[i], byte)][threadIdx.x]++;
  }
  __syncthreads();
  if (t < T)
    for (int d = 0; d < 256; d++) hist[(size_t)d * T + t] = cnt[d][threadIdx.x];
}
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const Fr* __restrict__ keys, const uint32_t* __restrict__ idx_in, uint32_t n,
                                                           uint32_t byte, uint32_t T, const uint32_t* __restrict__ offs,
                                                           uint32_t* __restrict__ idx_out) {
  __shared__ uint32_t pos[256][RS_THREADS];
  const uint32_t t = blockIdx.x * RS_THREADS + threadIdx.x;
  if (t >= T) return;
  for (int d = 0; d < 256; d++) pos[d][threadIdx.x] = offs[(size_t)d * T + t];
  const uint32_t b = t * RS_CHUNK, e = min(b + RS_CHUNK, n);
  for (uint32_t i = b; i < e; i++) {
    uint32_t id = idx_in[i];
    uint32_t d = key_byte(keys, id, byte);
    idx_out[pos[d][threadIdx.x]++] = id;
  }
}

// sortedness check of the full 256-bit keys + unique flags
__global__ void k_check_unique(const Fr* __restrict__ keys, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ flag,
                               uint32_t* __restrict__ unsorted) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) {
    flag[0] = 1;
    return;
  }
  int c = cmp256(ldf(keys + idx[i - 1]), ldf(keys + idx[i]));
  flag[i] = c != 0;
  if (c > 0) atomicOr(unsorted, 1u);
}
// U[upos[i]] = key, ustart[upos[i]] = i at unique positions
__global__ void k_collect_unique(const Fr* __restrict__ keys, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ flag,
                                 const uint32_t* __restrict__ upos, uint32_t n, Fr* __restrict__ U, uint32_t* __restrict__ ustart) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) {
    stf(U + upos[i], ldf(keys + idx[i]));
    ustart[upos[i]] = i;
  }
  if (i == 0) ustart[upos[n]] = n;
}

__device__ __forceinline__ int find_rank(const Fr* U, uint32_t D, const Fr& a) {
  uint32_t lo = 0, hi = D;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    int c = cmp256(ldf(U + mid), a);
    if (c < 0) lo = mid + 1; else hi = mid;
  }
  if (lo < D && cmp256(ldf(U + lo), a) == 0) return (int)lo;
  return -1;
}
__global__ void k_rank_inputs(const Fr* __restrict__ A, uint32_t n, const Fr* __restrict__ U, const uint32_t* __restrict__ Dptr,
                              uint32_t* __restrict__ rank, uint32_t* __restrict__ cntA, uint32_t* __restrict__ err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = find_rank(U, *Dptr, ldf(A + i));
  if (r < 0) {
    atomicOr(err, 1u);
    rank[i] = 0xffffffffu;
    return;
  }
  rank[i] = (uint32_t)r;
  atomicAdd(&cntA[r], 1u);
}
// first[r] = cntA[r] > 0 ; left[r] = cntT[r] - first[r]
__global__ void k_first_left(const uint32_t* __restrict__ cntA, const uint32_t* __restrict__ ustart, const uint32_t* __restrict__ Dptr,
                             uint32_t cap, uint32_t* __restrict__ first, uint32_t* __restrict__ left) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cap) return;
  if (r >= *Dptr) {
    first[r] = 0;
    left[r] = 0;
    return;
  }
  uint32_t f = cntA[r] > 0;
  first[r] = f;
  left[r] = (ustart[r + 1] - ustart[r]) - f;
}
// place every input: A'[slot] = U[r]; S'[slot] = own value at a first occurrence, else the left-over
// table copy number (R - 1 - repeated_index) in ascending order
__global__ void k_place(const uint32_t* __restrict__ rank, uint32_t n, const Fr* __restrict__ U, const uint32_t* __restrict__ Dptr,
                        const uint32_t* __restrict__ startA, uint32_t* __restrict__ cursorA, const uint32_t* __restrict__ dcount_excl,
                        const uint32_t* __restrict__ lstart, Fr* __restrict__ pa, Fr* __restrict__ ps) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r = rank[i];
  if (r == 0xffffffffu) return;
  const uint32_t D = *Dptr;
  uint32_t k = atomicAdd(&cursorA[r], 1u);
  uint32_t slot = startA[r] + k;
  Fr val = ldf(U + r);
  stf(pa + slot, fp_to_mont(val));
  if (k == 0) {
    stf(ps + slot, fp_to_mont(val));
    return;
  }
  const uint32_t distinct_total = dcount_excl[D];
  const uint32_t R = n - distinct_total;
  const uint32_t rr = slot - dcount_excl[r + 1];       // first-occurrence rows at or before `slot`
  const uint32_t j = R - 1 - rr;
  // last r2 with lstart[r2] <= j
  uint32_t lo = 0, hi = D;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (lstart[mid] <= j) lo = mid; else hi = mid;
  }
  stf(ps + slot, fp_to_mont(ldf(U + lo)));
}

inline uint32_t nb(uint32_t n, uint32_t t = 256) { return (n + t - 1) / t; }

}  // namespace

size_t lookup_workspace_bytes(uint32_t n) {
  size_t T = (n + RS_CHUNK - 1) / RS_CHUNK;
  size_t words = 0;
  words += 2 * (size_t)n;              // idx ping-pong
  words += 2 * 256 * T + 2;            // hist, offs
  words += 9 * ((size_t)n + 8);        // flag, upos, ustart, rank, cntA, startA, cursorA, first/left reuse, lstart, dcount
  return words * 4 + 3 * (size_t)n * sizeof(Fr) + 4096;
}

int lookup_permute(const Fr* a_mont, const Fr* s_mont, uint32_t usable, Fr* pa, Fr* ps, uint8_t* ws, uint32_t* status_dev,
                   bool full_sort, cudaStream_t st, LaunchCounter lc) {
  const uint32_t n = usable;
  const uint32_t T = (n + RS_CHUNK - 1) / RS_CHUNK;
  // carve the workspace
  Fr* Acan = (Fr*)ws;
  Fr* Tcan = Acan + n;
  Fr* U = Tcan + n;
  uint32_t* w = (uint32_t*)(U + n);
  uint32_t* idx0 = w; w += n;
  uint32_t* idx1 = w; w += n;
  uint32_t* hist = w; w += 256 * (size_t)T + 1;
  uint32_t* offs = w; w += 256 * (size_t)T + 1;
  uint32_t* flag = w; w += n + 8;
  uint32_t* upos = w; w += n + 8;
  uint32_t* ustart = w; w += n + 8;
  uint32_t* rank = w; w += n + 8;
  uint32_t* cntA = w; w += n + 8;
  uint32_t* startA = w; w += n + 8;
  uint32_t* cursorA = w; w += n + 8;
  uint32_t* first = w; w += n + 8;
  uint32_t* left = w; w += n + 8;
  // lstart / dcount reuse hist / offs (dead after the sort): both hold >= n + 1 words when T >= 1
  uint32_t* lstart = hist;
  uint32_t* dcount = offs;

  k_canonical<<<nb(n), 256, 0, st>>>(a_mont, Acan, n); lc++;
  k_canonical<<<nb(n), 256, 0, st>>>(s_mont, Tcan, n); lc++;
  k_iota<<<nb(n), 256, 0, st>>>(idx0, n); lc++;
  uint32_t* in = idx0;
  uint32_t* out = idx1;
  for (uint32_t byte = full_sort ? 0 : 24; byte < 32; byte++) {
    k_rs_count<<<nb(T, RS_THREADS), RS_THREADS, 0, st>>>(Tcan, in, n, byte, T, hist); lc++;
    k_scan_excl_u32<<<1, 1024, 0, st>>>(hist, 256 * T, offs); lc++;
    k_rs_scatter<<<nb(T, RS_THREADS), RS_THREADS, 0, st>>>(Tcan, in, n, byte, T, offs, out); lc++;
    uint32_t* tmp = in; in = out; out = tmp;
  }
  // status_dev[0] = unsorted flag (top-64-bit sort insufficient), status_dev[1] = input not in table
  k_check_unique<<<nb(n), 256, 0, st>>>(Tcan, in, n, flag, status_dev); lc++;
  k_scan_excl_u32<<<1, 1024, 0, st>>>(flag, n, upos); lc++;
  k_collect_unique<<<nb(n), 256, 0, st>>>(Tcan, in, flag, upos, n, U, ustart); lc++;
  const uint32_t* Dptr = upos + n;   // number of distinct table values
  cudaMemsetAsync(cntA, 0, (size_t)(n + 8) * 4, st);
  cudaMemsetAsync(cursorA, 0, (size_t)(n + 8) * 4, st);
  k_rank_inputs<<<nb(n), 256, 0, st>>>(Acan, n, U, Dptr, rank, cntA, status_dev + 1); lc++;
  k_scan_excl_u32<<<1, 1024, 0, st>>>(cntA, n, startA); lc++;
  k_first_left<<<nb(n), 256, 0, st>>>(cntA, ustart, Dptr, n, first, left); lc++;
  k_scan_excl_u32<<<1, 1024, 0, st>>>(left, n, lstart); lc++;
  k_scan_excl_u32<<<1, 1024, 0, st>>>(first, n, dcount); lc++;
  k_place<<<nb(n), 256, 0, st>>>(rank, n, U, Dptr, startA, cursorA, dcount, lstart, pa, ps); lc++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace zg
