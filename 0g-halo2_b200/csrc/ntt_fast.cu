// Multi-pass Cooley-Tukey NTT over BN254 Fr with register-resident radix-8 rounds (sm_100a), log2(N) >= 12.
//
// Same contract as ntt.cu (natural order in, natural order out; replaces halo2_proofs v2023_04_20 `best_fft` and the
// EvaluationDomain transforms reached from /root/reference/src/wnn.rs:242-259), different schedule:
//   * N = N_1 * ... * N_P with N_p = 2^S_p, 6 <= S_p <= 9 (P = 2 up to 2^18, 3 up to 2^27).  Pass p runs, for every
//     setting of the other digits, one N_p-point DFT over digit p and multiplies by the inter-pass twiddle
//     w_M^(k_p * i_low) (M = N_p * ... * N_P).  All twiddles INSIDE a pass are powers of w_{N_p}: they live in an
//     8-KB shared-memory table, not in a per-butterfly stream from L2 / DRAM (ncu on the stage-wise kernel:
//     long_scoreboard 3.1 per issue, lts hit 20 %, profiles/r01_ncu_full_ntt_pass_small_proof.txt).
//   * a CTA owns a tile of 2^S rows x 4 columns; a thread holds 8 elements (64 registers) and runs three radix-2 stages
//     on them per round (12 butterflies, 4 independent products in flight per stage), exchanging through shared
//     memory between rounds: 3 barriers per pass instead of one per stage.  The last round of every pass needs only
//     powers of w_8, so a 2^18 transform costs 8.25 products per element against 9 for stage-wise radix 2.
//   * global accesses are 128-byte runs in every pass: column passes move 4 adjacent columns; the last pass reads
//     4 whole rows whose leading output digit is adjacent and writes 4-element runs of the natural-order result.
//   * zero padding, per-element input scaling (coset shifts: a table of g^i, or the zeta pattern of halo2's extended
//     coset), output scaling (1/N, optionally per element) and truncation are fused into the first / last pass;
//     one launch covers a whole batch of polynomials and, for the prover's 3 x 2n extended domain, the three cosets.
// The arithmetic is integer-pipe bound: the kernel is judged against the Montgomery-product rate.
// The rounds are unrolled over a thread's 8 register-resident elements; with the multiplier inlined one pass was 212 KB of
// SASS (cuobjdump) and ran slower than the stage-wise kernel (2^20: 0.277 vs 0.236 ms) on instruction fetch.  One shared
// copy of the multiplier and ONE round body executed in a run-time loop keep the kernel near 30 KB.
#define ZG_FP_MUL_NOINLINE 1
#include <atomic>
#include "ntt.cuh"

namespace zg {

namespace {

constexpr int NF_LOGC = 2, NF_C = 4;

__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Fr ld_fr_plain(const Fr* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fr_plain(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// two planes of 16-byte halves (as in ntt.cu): 128-bit shared accesses, consecutive elements 16 B apart per plane
struct Planes {
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

// One round: the thread holds the 8 elements whose row index j differs in bits pc+2, pc+1, pc (slot u = those bits) and
// runs the decimation-in-frequency stages of the ACTIVE ones among them, most significant first.  Stage at bit s pairs
// rows j and j + 2^s and multiplies the difference by w_R^((j mod 2^s) << (S-1-s)).  pc / active are run-time values so
// that the three rounds of a pass share one copy of this code.
template <int S>
__device__ __forceinline__ void nf_round(Fr (&x)[8], uint32_t jbase, uint32_t pc, uint32_t active, const Planes& tw) {
#pragma unroll
  for (int q = 2; q >= 0; q--) {
    if (!((active >> q) & 1u)) continue;
    const uint32_t s = pc + q;
    const uint32_t smask = (1u << s) - 1u, sh = S - 1 - s;
#pragma unroll
    for (int u0 = 0; u0 < 8; u0++) {
      if ((u0 >> q) & 1) continue;
      const int u1 = u0 | (1 << q);
      const uint32_t jl = (jbase | ((uint32_t)u0 << pc)) & smask;
      Fr a = x[u0], b = x[u1];
      x[u0] = fp_add(a, b);
      Fr d = fp_sub(a, b);
      if (jl) d = fp_mul(d, tw.get(jl << sh));        // w^0 in the last round for half of the pairs, rarely elsewhere
      x[u1] = d;
    }
  }
}

// pc of each round (packed 4 bits each, first round lowest) and the active-stage masks (3 bits each)
template <int S> struct NfRounds;
template <> struct NfRounds<9> { static constexpr uint32_t N = 3, PC = 6 | (3 << 4) | (0 << 8), ACT = 7 | (7 << 4) | (7 << 8); };
template <> struct NfRounds<8> { static constexpr uint32_t N = 3, PC = 5 | (2 << 4) | (0 << 8), ACT = 7 | (7 << 4) | (3 << 8); };
template <> struct NfRounds<7> { static constexpr uint32_t N = 3, PC = 4 | (1 << 4) | (0 << 8), ACT = 7 | (7 << 4) | (1 << 8); };
template <> struct NfRounds<6> { static constexpr uint32_t N = 2, PC = 3 | (0 << 4), ACT = 7 | (7 << 4); };

}  // namespace

template <int S>
__global__ void __launch_bounds__((1 << S) * NF_C / 8, (S == 9 ? 2 : S == 8 ? 4 : S == 7 ? 8 : 16)) ntt_fast_pass_kernel(NttFastArgs A) {
  constexpr uint32_t R = 1u << S, T = R * NF_C / 8, TOTAL = R * NF_C;
  extern __shared__ uint4 nf_smem[];
  Planes sm{nf_smem, nf_smem + TOTAL};
  Planes tw{nf_smem + 2 * TOTAL, nf_smem + 2 * TOTAL + R / 2};
  const uint32_t tid = threadIdx.x, tile = blockIdx.x;
  const uint32_t poly = blockIdx.y / A.cosets, cz = blockIdx.y - poly * A.cosets;
  const Fr* in = A.in + (size_t)poly * A.in_stride + (size_t)cz * A.in_coset_stride;
  Fr* out = A.out + (size_t)poly * A.out_stride + (size_t)cz * A.out_coset_stride;
  const uint32_t logn = A.logn, lo = A.lo, hi_bits = A.hi_bits;

  // tile geometry
  uint32_t base = 0, low0 = 0, hi = 0, rest_hi = 0, kap4 = 0, hi_rest_bits = 0;
  if (!A.last) {
    const uint32_t lowbits = lo - NF_LOGC;
    low0 = (tile & ((1u << lowbits) - 1u)) << NF_LOGC;
    hi = tile >> lowbits;
    base = (hi << (lo + S)) | low0;
  } else {
    hi_rest_bits = hi_bits - A.dig[0];
    rest_hi = tile & ((1u << hi_rest_bits) - 1u);
    kap4 = (tile >> hi_rest_bits) << NF_LOGC;        // first of the 4 adjacent leading digits
  }
  // w_R^x, x < R/2, from the flat table w_N^e
  for (uint32_t x = tid; x < R / 2; x += T) tw.put(x, ldg_fr(A.flat + ((size_t)x << (logn - S))));
  // tile -> shared memory (coalesced along whatever is contiguous in global memory)
  for (uint32_t e = tid; e < TOTAL; e += T) {
    uint32_t j, col, idx;
    if (!A.last) {
      col = e & (NF_C - 1);
      j = e >> NF_LOGC;
      idx = base + (j << lo) + col;
    } else {
      j = e & (R - 1);
      col = e >> S;
      idx = ((((kap4 + col) << hi_rest_bits) | rest_hi) << S) + j;
    }
    Fr v;
    if (idx < A.n_in) {
      v = A.first ? ldg_fr(in + idx) : ld_fr_plain(in + idx);
      if (A.first) {
        if (A.in_table) v = fp_mul(v, ldg_fr(A.in_table + (size_t)cz * A.in_table_stride + idx));
        else if (A.flags & NTT_IN_COSET) {
          const uint32_t m = idx % 3;
          if (m) v = fp_mul(v, A.in_scale[m]);
        }
      }
    } else {
      v = fp_zero<FrParams>();
    }
    sm.put((j << NF_LOGC) | col, v);
  }
  __syncthreads();

  const uint32_t col = tid & (NF_C - 1), rest = tid >> NF_LOGC;
  Fr x[8];
  using RD = NfRounds<S>;
#pragma unroll 1
  for (uint32_t r = 0; r < RD::N; r++) {
    const uint32_t pc = (RD::PC >> (4 * r)) & 15u, active = (RD::ACT >> (4 * r)) & 15u;
    const uint32_t jbase = ((rest >> pc) << (pc + 3)) | (rest & ((1u << pc) - 1u));   // thread bits around the register bits
#pragma unroll
    for (int u = 0; u < 8; u++) x[u] = sm.get(((jbase | ((uint32_t)u << pc)) << NF_LOGC) | col);
    nf_round<S>(x, jbase, pc, active, tw);
    if (r + 1 < RD::N) {
#pragma unroll
      for (int u = 0; u < 8; u++) sm.put(((jbase | ((uint32_t)u << pc)) << NF_LOGC) | col, x[u]);
      __syncthreads();
    }
  }

  // registers -> global: slot u holds row j = (rest << 3) | u, i.e. output digit k = bitrev_S(j)
  const uint32_t krest = __brev(rest) >> (32 - (S - 3));       // bitrev of the S-3 thread bits = low bits of k
  if (!A.last) {
    const uint32_t low = low0 + col;
    const uint32_t half = 1u << (logn - 1);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint32_t k = ((__brev((uint32_t)u) >> 29) << (S - 3)) | krest;
      uint32_t e = (k * low) << hi_bits;                        // exponent of w_N; k * low < 2^(S + lo)
      Fr v = x[u];
      if (e) {
        const bool neg = e >= half;
        Fr w = ldg_fr(A.flat + (neg ? e - half : e));
        v = fp_mul(v, w);
        if (neg) v = fp_neg(v);
      }
      st_fr_plain(out + ((hi << (lo + S)) | (k << lo) | low), v);
    }
  } else {
    // natural-order position: leading digit first, then the other upper digits, this pass's digit on top
    const uint32_t d0 = kap4 + col;
    uint32_t posl = d0, shift = A.dig[0], rem = rest_hi, rem_bits = hi_rest_bits;
    for (uint32_t t = 1; t < A.ndig; t++) {
      rem_bits -= A.dig[t];
      posl |= (rem >> rem_bits) << shift;
      rem &= (1u << rem_bits) - 1u;
      shift += A.dig[t];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint32_t k = ((__brev((uint32_t)u) >> 29) << (S - 3)) | krest;
      const uint32_t pos = posl | (k << hi_bits);
      if (pos >= A.n_out) continue;
      Fr v = x[u];
      if (A.out_table) v = fp_mul(v, ldg_fr(A.out_table + (size_t)cz * A.out_table_stride + pos));
      else if (A.flags & NTT_OUT_SCALE) v = fp_mul(v, A.out_scale[(A.flags & NTT_OUT_MOD3) ? pos % 3 : 0]);
      st_fr_plain(out + pos, v);
    }
  }
}

// flat[e] = w^e for e < N/2 is built by ntt_build_twiddles (ntt.cu)

bool ntt_fast_supported(uint32_t logn) { return logn >= 12 && logn <= 28; }
// tiles of 2^S x 4 elements: below about two CTAs per SM the stage-wise kernel (smaller tiles, more passes) fills the GPU better
bool ntt_fast_pays(uint32_t logn, uint32_t transforms) {
  const uint32_t npass = (logn + 8) / 9, s0 = (logn + npass - 1) / npass;
  return ((uint64_t)transforms << (logn - s0 - NF_LOGC)) >= 296;
}

template <int S>
static void nf_launch(const NttFastArgs& A, dim3 grid, cudaStream_t st) {
  constexpr uint32_t R = 1u << S;
  const size_t smem = (size_t)(2 * R * NF_C + R) * sizeof(uint4);
  static std::atomic<uint64_t> attr_devices{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_devices.load(std::memory_order_acquire) >> (dev & 63) & 1)) {
    cudaFuncSetAttribute(ntt_fast_pass_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  ntt_fast_pass_kernel<S><<<grid, R * NF_C / 8, smem, st>>>(A);
}

cudaError_t ntt_fast_run(const NttPlan& P, cudaStream_t st, uint64_t* nl) {
  const uint32_t logn = P.logn;
  const uint32_t npass = (logn + 8) / 9;
  uint32_t S[4] = {0, 0, 0, 0};
  for (uint32_t p = 0, left = logn; p < npass; p++) {       // as even as possible, larger digits first
    S[p] = (left + (npass - p) - 1) / (npass - p);
    left -= S[p];
  }
  const uint32_t cosets = P.cosets ? P.cosets : 1;
  uint32_t consumed = 0;
  for (uint32_t p = 0; p < npass; p++) {
    NttFastArgs A{};
    const bool first = p == 0, last = p == npass - 1;
    A.flat = P.flat;
    A.logn = logn;
    A.lo = logn - consumed - S[p];
    A.hi_bits = consumed;
    A.first = first;
    A.last = last;
    A.cosets = cosets;
    A.n_in = first ? P.n_in : (1u << logn);
    A.n_out = last ? P.n_out : (1u << logn);
    A.flags = (first ? (P.flags & NTT_IN_COSET) : 0) | (last ? (P.flags & (NTT_OUT_SCALE | NTT_OUT_MOD3)) : 0);
    for (int i = 0; i < 3; i++) { A.in_scale[i] = P.in_scale[i]; A.out_scale[i] = P.out_scale[i]; }
    A.in_table = first ? P.in_table : nullptr;
    A.in_table_stride = P.in_table_stride;
    A.out_table = last ? P.out_table : nullptr;
    A.out_table_stride = P.out_table_stride;
    // routing: first pass reads the caller's input; passes in between run in place in tmp (one slot per polynomial and
    // coset); the last pass permutes into the caller's output
    if (first) {
      A.in = P.in; A.in_stride = P.in_stride; A.in_coset_stride = P.in_coset_stride;
    } else {
      A.in = P.tmp; A.in_stride = P.tmp_stride * cosets; A.in_coset_stride = P.tmp_stride;
    }
    if (last) {
      A.out = P.out; A.out_stride = P.out_stride; A.out_coset_stride = P.out_coset_stride;
    } else {
      A.out = P.tmp; A.out_stride = P.tmp_stride * cosets; A.out_coset_stride = P.tmp_stride;
    }
    A.ndig = p;
    for (uint32_t t = 0; t < p; t++) A.dig[t] = S[t];
    const uint32_t tiles = 1u << (logn - S[p] - NF_LOGC);
    dim3 grid(tiles, P.batch * cosets);
    switch (S[p]) {
      case 6: nf_launch<6>(A, grid, st); break;
      case 7: nf_launch<7>(A, grid, st); break;
      case 8: nf_launch<8>(A, grid, st); break;
      default: nf_launch<9>(A, grid, st); break;
    }
    if (nl) ++*nl;
    consumed += S[p];
  }
  return cudaGetLastError();
}

}  // namespace zg
