// Radix-2 multi-pass NTT over BN254 Fr for sm_100a.
//
// Replaces halo2_proofs (tag v2023_04_20, un-vendored; pinned by /root/reference/Cargo.toml:21-25)
// `arithmetic::best_fft` and the EvaluationDomain transforms built on it
// (`lagrange_to_coeff`, `coeff_to_extended`, `extended_to_coeff`), reached from
// /root/reference/src/wnn.rs:242-259 (create_proof) and :226-228 (keygen).
//
// Layout and algorithm (B200-first, not a translation of the upstream bit-reverse + recursive
// butterflies):
//   * natural order in -> natural order out, as best_fft.
//   * decimation-in-frequency stages are grouped into passes of S index bits; one CTA owns a
//     tile of 2^S rows x C columns in shared memory (C contiguous elements = C*32 B per row, so
//     every global access is a full 128-B line for C = 4), runs its S stages there, and
//     writes back in place.  The last pass owns the low S bits, pulls C tiles that differ in
//     the TOP index bits, and stores bit-reversed, which makes its stores C*32-B contiguous.
//   * twiddles come from a stage-major table T_t[j] = w_{2^(t+1)}^j (offset 2^t - 1) so that
//     lanes that walk consecutive columns read consecutive table entries.
//   * 254-bit Montgomery butterflies make this kernel integer-pipe bound (about 10 mulmods per
//     element at 2^20 against 64 B of traffic per pass), see DESIGN.md: ncu inside the k = 17 proof shows
//     sm__pipe_fmaheavy_cycles_active 67-82 % for the batched passes (profiles/r02_ncu_full_ntt_proof_large.txt).
//   * one launch covers a batch of polynomials and, per polynomial, several COSETS that differ only in a per-element
//     input table (coset c multiplies coefficient i by g_c^i) or output table: the prover's internal extended domain
//     is three cosets of the 2n-point subgroup (extdomain.cuh), so coefficient -> extended is one launch per pass for
//     all columns of a round and all three cosets; a rank that owns a subset of the cosets passes (first, step).
#include <atomic>
#include <cstdlib>
#include "ntt.cuh"

namespace zg {

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// ---- twiddle tables --------------------------------------------------------------------
// flat[i] = w^i for i < half; each thread seeds with a pow and walks CHUNK entries.
static constexpr int TW_CHUNK = 32;
__global__ void ntt_twiddle_flat_kernel(Fr* flat, Fr w, uint32_t half) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t start = (uint64_t)t * TW_CHUNK;
  if (start >= half) return;
  Fr cur = fp_pow_u64(w, start);
  for (int i = 0; i < TW_CHUNK && start + i < half; i++) {
    st_fr(flat + start + i, cur);
    cur = fp_mul(cur, w);
  }
}
// stage-major gather: T[(1<<t)-1+j] = flat[j << (logn-1-t)]
__global__ void ntt_twiddle_stage_kernel(Fr* tab, const Fr* flat, uint32_t logn) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. n-2
  uint32_t n = 1u << logn;
  if (e >= n - 1) return;
  uint32_t t = 31 - __clz(e + 1);
  uint32_t j = e + 1 - (1u << t);
  st_fr(tab + e, ld_fr(flat + ((size_t)j << (logn - 1 - t))));
}

struct SmTile {
  uint4* lo;
  uint4* hi;
  __device__ __forceinline__ Fr get(uint32_t e) const {
    uint4 a = lo[e], b = hi[e];
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ void put(uint32_t e, const Fr& r) const {
    lo[e] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[e] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
};

// ---- one pass ----------------------------------------------------------------------------
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(NttPassArgs A) {
  // The tile is kept as two planes of 16-byte halves (low limbs | high limbs): consecutive elements are 16 B apart in
  // each plane, so the 128-bit shared accesses of a warp hit 32 distinct banks per quarter-warp.  (Whole 32-byte
  // elements gave a 2-way conflict on every access: ncu l1tex__data_bank_conflicts 4.2 M of 8.4 M wavefronts.)
  extern __shared__ uint4 ntt_smem_raw[];
  SmTile sm{ntt_smem_raw, ntt_smem_raw + ((1u << A.S) << A.logc)};
  const uint32_t S = A.S, logc = A.logc, lo = A.lo, logn = A.logn;
  const uint32_t R = 1u << S, C = 1u << logc;
  const uint32_t poly = blockIdx.y / A.cosets, cz = blockIdx.y - poly * A.cosets;
  const uint32_t ctab = A.tab_cid_base + cz * A.tab_cid_step;
  const Fr* in = A.in + (size_t)poly * A.in_stride + (size_t)(A.in_cid_base + cz * A.in_cid_step) * A.in_coset_stride;
  Fr* out = A.out + (size_t)poly * A.out_stride + (size_t)(A.out_cid_base + cz * A.out_cid_step) * A.out_coset_stride;
  const uint32_t tile = blockIdx.x;
  const uint32_t tid = threadIdx.x;
  const uint32_t total = R << logc;

  uint32_t base, mid = 0;
  if (!A.contiguous) {
    const uint32_t midbits = lo - logc;
    mid = tile & ((1u << midbits) - 1);
    const uint32_t top = tile >> midbits;
    base = (top << (lo + S)) | (mid << logc);
    for (uint32_t e = tid; e < total; e += NTT_THREADS) {
      uint32_t c = e & (C - 1), j = e >> logc;
      uint32_t idx = base | (j << lo) | c;
      Fr v;
      if (idx < A.n_in) {
        v = ld_fr(in + idx);
        if (A.in_table) v = fp_mul(v, ld_fr(A.in_table + (size_t)ctab * A.in_table_stride + idx));
        else if (A.flags & NTT_IN_COSET) {
          uint32_t m = idx % 3;
          if (m) v = fp_mul(v, A.in_scale[m]);
        }
      } else {
        v = fp_zero<FrParams>();
      }
      sm.put(e, v);  // e = j*C + c
    }
  } else {
    base = tile << S;  // `rest` field sits above the S active bits
    for (uint32_t e = tid; e < total; e += NTT_THREADS) {
      uint32_t j = e & (R - 1), c = e >> S;
      uint32_t idx = (c << (logn - logc)) | base | j;
      if (logc == 0) idx = base | j;
      Fr v;
      if (idx < A.n_in) {
        v = ld_fr(in + idx);
        if (A.in_table) v = fp_mul(v, ld_fr(A.in_table + (size_t)ctab * A.in_table_stride + idx));
        else if (A.flags & NTT_IN_COSET) {
          uint32_t m = idx % 3;
          if (m) v = fp_mul(v, A.in_scale[m]);
        }
      } else {
        v = fp_zero<FrParams>();
      }
      sm.put(e, v);  // e = c*R + j
    }
  }
  __syncthreads();

  const uint32_t nbf = total >> 1;
  for (int m = (int)S - 1; m >= 0; m--) {
    const uint32_t half = 1u << m;
    const uint32_t t = lo + m;
    const Fr* T = A.tw + ((1u << t) - 1);
    for (uint32_t b = tid; b < nbf; b += NTT_THREADS) {
      uint32_t c, jb;
      if (!A.contiguous) {
        c = b & (C - 1);
        jb = b >> logc;
      } else {
        jb = b & ((R >> 1) - 1);
        c = b >> (S - 1);
      }
      const uint32_t jl = jb & (half - 1);
      const uint32_t ja = ((jb >> m) << (m + 1)) | jl;
      uint32_t ea, eb, widx;
      if (!A.contiguous) {
        ea = (ja << logc) | c;
        eb = ea + (half << logc);
        widx = (jl << lo) | (mid << logc) | c;
      } else {
        ea = (c << S) | ja;
        eb = ea + half;
        widx = jl;
      }
      Fr x = sm.get(ea), y = sm.get(eb);
      Fr s = fp_add(x, y);
      Fr d = fp_sub(x, y);
      if (widx != 0) d = fp_mul(d, ld_fr(T + widx));
      sm.put(ea, s);
      sm.put(eb, d);
    }
    __syncthreads();
  }

  if (!A.contiguous) {
    for (uint32_t e = tid; e < total; e += NTT_THREADS) {
      uint32_t c = e & (C - 1), j = e >> logc;
      uint32_t idx = base | (j << lo) | c;
      st_fr(out + idx, sm.get(e));
    }
  } else {
    const uint32_t restbits = logn - S - logc;
    const uint32_t rest_rev = restbits ? (__brev(tile) >> (32 - restbits)) : 0;
    for (uint32_t e = tid; e < total; e += NTT_THREADS) {
      uint32_t cr = e & (C - 1), jr = e >> logc;
      uint32_t j = __brev(jr) >> (32 - S);
      uint32_t c = logc ? (__brev(cr) >> (32 - logc)) : 0;
      uint32_t pos = (jr << (logn - S)) | (rest_rev << logc) | cr;
      if (pos >= A.n_out) continue;
      Fr v = sm.get((c << S) | j);
      if (A.out_table) v = fp_mul(v, ld_fr(A.out_table + (size_t)ctab * A.out_table_stride + pos));
      else if (A.flags & NTT_OUT_SCALE) v = fp_mul(v, A.out_scale[(A.flags & NTT_OUT_MOD3) ? pos % 3 : 0]);
      st_fr(out + pos, v);
    }
  }
}

// ---- host-side planning ----------------------------------------------------------------
cudaError_t ntt_build_twiddles(Fr* tab, Fr* scratch_flat, const Fr& w, uint32_t logn,
                               cudaStream_t stream) {
  uint32_t half = 1u << (logn - 1);
  uint32_t nthreads = (half + TW_CHUNK - 1) / TW_CHUNK;
  ntt_twiddle_flat_kernel<<<(nthreads + 127) / 128, 128, 0, stream>>>(scratch_flat, w, half);
  uint32_t n1 = (1u << logn) - 1;
  ntt_twiddle_stage_kernel<<<(n1 + 255) / 256, 256, 0, stream>>>(tab, scratch_flat, logn);
  return cudaGetLastError();
}

// Passes are planned top-down; every strided pass keeps lo >= NTT_LOGC.
cudaError_t ntt_run(const NttPlan& P, cudaStream_t stream, uint64_t* nl) {
  const uint32_t logn = P.logn;
  uint32_t npass = (logn + NTT_MAX_S - 1) / NTT_MAX_S;
  if (npass == 0) npass = 1;
  // a lone small transform would run on a handful of CTAs: trade one more pass for a grid that covers the SMs
  while (npass < 4 && logn >= 5 * (npass + 1)) {
    uint32_t S0 = (logn + npass - 1) / npass;
    uint64_t ctas = ((uint64_t)1 << (logn - S0 - (logn - S0 >= NTT_LOGC ? NTT_LOGC : logn - S0))) * P.batch * (P.cosets ? P.cosets : 1);
    if (ctas >= 148) break;
    npass++;
  }
  uint32_t bits_left = logn;
  uint32_t hi = logn;
  // the attribute is per device; one process may drive several devices / host threads
  static std::atomic<uint64_t> attr_devices{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_devices.load(std::memory_order_acquire) >> (dev & 63) & 1)) {
    cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((sizeof(Fr) << NTT_MAX_S) << NTT_LOGC));
    attr_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  for (uint32_t p = 0; p < npass; p++) {
    uint32_t S = (bits_left + (npass - p) - 1) / (npass - p);
    NttPassArgs A;
    A.tw = P.tw;
    A.logn = logn;
    A.S = S;
    A.lo = hi - S;
    A.flags = 0;
    A.n_in = 1u << logn;
    A.n_out = 1u << logn;
    const bool first = (p == 0), last = (p == npass - 1);
    // buffer routing: first pass reads `in`; last pass writes `out`; between them `tmp`
    // (the last pass permutes across tiles, so it can never run in place unless it is alone).
    const uint32_t cosets = P.cosets ? P.cosets : 1;
    A.cosets = cosets;
    A.in_table = first ? P.in_table : nullptr;
    A.in_table_stride = P.in_table_stride;
    A.out_table = nullptr;
    A.out_table_stride = P.out_table_stride;
    A.in = first ? P.in : P.tmp;
    A.in_stride = first ? P.in_stride : P.tmp_stride * cosets;
    A.in_coset_stride = first ? P.in_coset_stride : P.tmp_stride;
    A.tab_cid_base = P.coset_first; A.tab_cid_step = P.coset_step;
    A.in_cid_base = first ? P.coset_first : 0; A.in_cid_step = first ? P.coset_step : 1;
    A.out_cid_base = last ? P.coset_first : 0; A.out_cid_step = last ? P.coset_step : 1;
    if (last) {
      A.out = P.out;
      A.out_stride = P.out_stride;
      A.out_coset_stride = P.out_coset_stride;
      A.out_table = P.out_table;
    } else {
      A.out = P.tmp;
      A.out_stride = P.tmp_stride * cosets;
      A.out_coset_stride = P.tmp_stride;
    }
    if (first) {
      A.n_in = P.n_in;
      if (P.flags & NTT_IN_COSET) {
        A.flags |= NTT_IN_COSET;
        for (int i = 0; i < 3; i++) A.in_scale[i] = P.in_scale[i];
      }
    }
    if (last) {
      A.contiguous = 1;
      A.logc = (logn - S >= NTT_LOGC) ? NTT_LOGC : (logn - S);
      A.n_out = P.n_out;
      if (P.flags & NTT_OUT_SCALE) {
        A.flags |= (P.flags & (NTT_OUT_SCALE | NTT_OUT_MOD3));
        for (int i = 0; i < 3; i++) A.out_scale[i] = P.out_scale[i];
      }
    } else {
      A.contiguous = 0;
      A.logc = NTT_LOGC;
    }
    uint32_t tiles = 1u << (logn - S - A.logc);
    size_t smem = (sizeof(Fr) << S) << A.logc;
    dim3 grid(tiles, P.batch * cosets);
    ntt_pass_kernel<<<grid, NTT_THREADS, smem, stream>>>(A);
    if (nl) ++*nl;
    hi -= S;
    bits_left -= S;
  }
  return cudaGetLastError();
}

}  // namespace zg
