// Lookup argument: permute_expression_pair on the device.
//
// Replaces halo2_proofs v2023_04_20 (un-vendored; /root/reference/Cargo.toml:21-25)
// `plonk::lookup::prover::permute_expression_pair` -- upstream: single-threaded `sort()` of the
// compressed inputs, a BTreeMap multiset of the table, then "first occurrence takes its own value,
// remaining table values ascending go to the repeated rows from the highest row down"
// (SURVEY.md Appendix B.3).  Call site: create_proof, /root/reference/src/wnn.rs:242-259.
//
// B200-first restatement that yields the same two columns without sorting the inputs:
//   1. block-parallel LSD radix sort (8-bit digits, warp match_any ranking) of the canonical table
//      values only -- top 48 bits first, checked, full 256-bit passes as the fallback -- then the
//      unique values U with their multiplicities.  Tables that do not depend on theta (one table
//      expression) are sorted once per proving key and cached by the caller;
//   2. every input is ranked by binary search in U (an input that is not in the table raises the
//      constraint-system failure flag) and counted: cntA;
//   3. exclusive scans give the start of every value's run in A', the ascending list of left-over
//      table copies, and the index of every repeated row, from which each row of S' is a direct
//      look-up (no serial walk).
#include "lookup.cuh"
#include "scan.cuh"

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
__global__ void k_canonical_iota(const Fr* __restrict__ in, Fr* __restrict__ out, uint32_t* __restrict__ idx, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  stf(out + i, fp_from_mont(ldf(in + i)));
  idx[i] = i;
}

// ---- block-parallel LSD radix sort of an index permutation by one key byte ---------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 4;                       // rounds per CTA
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // keys per CTA

__device__ __forceinline__ uint32_t key_byte(const Fr* keys, uint32_t idx, uint32_t byte) {
  return (keys[idx].v[byte >> 2] >> ((byte & 3) * 8)) & 0xff;
}
// hist[d * nblocks + b]
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const Fr* __restrict__ keys, const uint32_t* __restrict__ idx, uint32_t n,
                                                        uint32_t byte, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    uint32_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[key_byte(keys, idx[i], byte)], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}
// stable scatter: position = offs[d][block] + (same-digit items earlier in this CTA)
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const Fr* __restrict__ keys, const uint32_t* __restrict__ idx_in, uint32_t n,
                                                           uint32_t byte, const uint32_t* __restrict__ offs,
                                                           uint32_t* __restrict__ idx_out) {
  __shared__ uint32_t cnt[RS_THREADS / 32][256];   // per-warp digit counts of the current round
  __shared__ uint32_t base_d[256];                 // running start of digit d for this CTA
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  base_d[tid] = offs[tid * gridDim.x + blockIdx.x];
  const uint32_t base = blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ITEMS; r++) {
    for (int d = lane; d < 256; d += 32) cnt[wid][d] = 0;
    __syncwarp();
    const uint32_t i = base + r * RS_THREADS + tid;
    const bool valid = i < n;
    uint32_t id = 0, d = 0xffffffffu;
    if (valid) {
      id = idx_in[i];
      d = key_byte(keys, id, byte);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1));
    if (valid && rank_in_warp == 0) cnt[wid][d] = __popc(peers);
    __syncthreads();
    // thread `tid` owns digit `tid`: exclusive prefix over the warps, then advance the CTA base
    {
      uint32_t run = base_d[tid];
#pragma unroll
      for (int w = 0; w < RS_THREADS / 32; w++) {
        uint32_t c = cnt[w][tid];
        cnt[w][tid] = run;
        run += c;
      }
      base_d[tid] = run;
    }
    __syncthreads();
    if (valid) idx_out[cnt[wid][d] + rank_in_warp] = id;
    __syncthreads();
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
// U[upos[i]] = key, ustart[upos[i]] = i at unique positions; ustart[D] = n; *D_out = D
__global__ void k_collect_unique(const Fr* __restrict__ keys, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ flag,
                                 const uint32_t* __restrict__ upos, uint32_t n, Fr* __restrict__ U, uint32_t* __restrict__ ustart,
                                 uint32_t* __restrict__ D_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) {
    stf(U + upos[i], ldf(keys + idx[i]));
    ustart[upos[i]] = i;
  }
  if (i == 0) {
    ustart[upos[n]] = n;
    *D_out = upos[n];
  }
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
// rank[i] of input i in U (Montgomery input converted on the fly) and the input histogram
__global__ void k_rank_inputs(const Fr* __restrict__ a_mont, uint32_t n, const Fr* __restrict__ U, const uint32_t* __restrict__ Dptr,
                              uint32_t* __restrict__ rank, uint32_t* __restrict__ cntA, uint32_t* __restrict__ err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = find_rank(U, *Dptr, fp_from_mont(ldf(a_mont + i)));
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
  Fr val = fp_to_mont(ldf(U + r));
  stf(pa + slot, val);
  if (k == 0) {
    stf(ps + slot, val);
    return;
  }
  const uint32_t distinct_total = dcount_excl[D];
  const uint32_t R = n - distinct_total;
  const uint32_t rr = slot - dcount_excl[r + 1];       // first-occurrence rows at or before `slot`
  const uint32_t j = R - 1 - rr;
  uint32_t lo = 0, hi = D;                              // last r2 with lstart[r2] <= j
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (lstart[mid] <= j) lo = mid; else hi = mid;
  }
  stf(ps + slot, fp_to_mont(ldf(U + lo)));
}

inline uint32_t nb(uint32_t n, uint32_t t = 256) { return (n + t - 1) / t; }

}  // namespace

size_t lookup_table_bytes(uint32_t n) { return (size_t)(n + 8) * sizeof(Fr) + (size_t)(n + 16) * 4; }

LookupTable lookup_table_carve(uint8_t* mem, uint32_t n) {
  LookupTable t;
  t.U = (Fr*)mem;
  t.ustart = (uint32_t*)(t.U + (n + 8));
  t.D = t.ustart + (n + 8);
  return t;
}

size_t lookup_workspace_bytes(uint32_t n) {
  size_t nblocks = (n + RS_TILE - 1) / RS_TILE;
  size_t words = 2 * (size_t)n + 2 * (256 * nblocks + 8) + 9 * ((size_t)n + 8);
  return words * 4 + (size_t)n * sizeof(Fr) + lookup_table_bytes(n) + 4096 + scan_scratch_words(n + 8, 3) * 4;
}
// tile sums of the multi-CTA scans: the last bytes of the workspace
static uint32_t* lookup_scan_scratch(uint8_t* ws, uint32_t n) {
  size_t off = lookup_workspace_bytes(n) - scan_scratch_words(n + 8, 3) * 4;
  return (uint32_t*)(ws + (off & ~(size_t)15));
}

LookupTable lookup_workspace_table(uint8_t* ws, uint32_t n) {
  // the scratch table lives at the end of the workspace
  size_t nblocks = (n + RS_TILE - 1) / RS_TILE;
  size_t words = 2 * (size_t)n + 2 * (256 * nblocks + 8) + 9 * ((size_t)n + 8);
  size_t off = (words * 4 + (size_t)n * sizeof(Fr) + 255) & ~(size_t)255;
  return lookup_table_carve(ws + off, n);
}

int lookup_sort_table(const Fr* s_mont, uint32_t usable, const LookupTable& out, uint8_t* ws, uint32_t* unsorted_flag_dev,
                      bool full_sort, cudaStream_t st, LaunchCounter lc) {
  const uint32_t n = usable;
  const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
  Fr* Tcan = (Fr*)ws;
  uint32_t* w = (uint32_t*)(Tcan + n);
  uint32_t* idx0 = w; w += n;
  uint32_t* idx1 = w; w += n;
  uint32_t* hist = w; w += 256 * (size_t)nblocks + 8;
  uint32_t* offs = w; w += 256 * (size_t)nblocks + 8;
  uint32_t* flag = w; w += n + 8;
  uint32_t* upos = w; w += n + 8;
  k_canonical_iota<<<nb(n), 256, 0, st>>>(s_mont, Tcan, idx0, n); lc++;
  uint32_t* in = idx0;
  uint32_t* o = idx1;
  for (uint32_t byte = full_sort ? 0 : 26; byte < 32; byte++) {   // partial sort: the top 48 bits, then checked
    k_rs_hist<<<nblocks, RS_THREADS, 0, st>>>(Tcan, in, n, byte, hist); lc++;
    {
      ScanJobs sj{};
      sj.in[0] = hist; sj.out[0] = offs;
      scan_excl_u32(sj, 1, 256 * nblocks, lookup_scan_scratch(ws, n), st, lc);
    }
    k_rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(Tcan, in, n, byte, offs, o); lc++;
    uint32_t* tmp = in; in = o; o = tmp;
  }
  k_check_unique<<<nb(n), 256, 0, st>>>(Tcan, in, n, flag, unsorted_flag_dev); lc++;
  {
    ScanJobs sj{};
    sj.in[0] = flag; sj.out[0] = upos;
    scan_excl_u32(sj, 1, n, lookup_scan_scratch(ws, n), st, lc);
  }
  k_collect_unique<<<nb(n), 256, 0, st>>>(Tcan, in, flag, upos, n, out.U, out.ustart, out.D); lc++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int lookup_permute(const Fr* a_mont, uint32_t usable, const LookupTable& tab, Fr* pa, Fr* ps, uint8_t* ws, uint32_t* missing_flag_dev,
                   cudaStream_t st, LaunchCounter lc) {
  const uint32_t n = usable;
  const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
  // bookkeeping arrays sit after the sort's own arrays (which are dead by now or unused for cached tables)
  uint32_t* w = (uint32_t*)((Fr*)ws + n);
  w += 2 * (size_t)n + 2 * (256 * (size_t)nblocks + 8) + 2 * ((size_t)n + 8);
  uint32_t* rank = w; w += n + 8;
  uint32_t* cntA = w; w += n + 8;        // cntA | cursorA contiguous: one memset
  uint32_t* cursorA = w; w += n + 8;
  uint32_t* startA = w; w += n + 8;
  uint32_t* first = w; w += n + 8;
  uint32_t* left = w; w += n + 8;
  uint32_t* lstart = w; w += n + 8;
  // dcount reuses the first index buffer of the sort
  uint32_t* dcount = (uint32_t*)((Fr*)ws + n);
  cudaMemsetAsync(cntA, 0, (size_t)2 * (n + 8) * 4, st);
  k_rank_inputs<<<nb(n), 256, 0, st>>>(a_mont, n, tab.U, tab.D, rank, cntA, missing_flag_dev); lc++;
  k_first_left<<<nb(n), 256, 0, st>>>(cntA, tab.ustart, tab.D, n, first, left); lc++;
  ScanJobs sj{};
  sj.in[0] = cntA; sj.out[0] = startA;
  sj.in[1] = left; sj.out[1] = lstart;
  sj.in[2] = first; sj.out[2] = dcount;
  scan_excl_u32(sj, 3, n, lookup_scan_scratch(ws, n), st, lc);
  k_place<<<nb(n), 256, 0, st>>>(rank, n, tab.U, tab.D, startA, cursorA, dcount, lstart, pa, ps); lc++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace zg
