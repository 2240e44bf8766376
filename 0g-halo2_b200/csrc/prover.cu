// Host orchestration of keygen and create_proof on one B200.
//
// Restates the stage order, Fiat-Shamir transcript and RNG draw order of halo2_proofs v2023_04_20
// (un-vendored; /root/reference/Cargo.toml:21-25) `plonk::{keygen_vk, keygen_pk, create_proof}` with
// `ProverGWC` and snark-verifier's `EvmTranscript`, as instantiated at /root/reference/src/wnn.rs:226-228
// and :242-259.  All column arithmetic runs in the kernels of ntt.cu,
// msm.cu, poly.cu, expr.cu and lookup.cu; the host only sequences rounds, hashes (keccak256) and
// normalises <= 8 commitments per round.  Witness synthesis is the caller's (BASELINE north_star).
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <memory>
#include "ctx.cuh"
#include "expr.cuh"
#include "extdomain.cuh"
#include "lookup.cuh"
#include "poly.cuh"

using namespace zg;

extern "C" int zg_msm_dev(zg_ctx*, int, const zg_fr*, size_t, size_t, size_t, zg_g1*);
extern "C" int zg_lagrange_to_coeff_dev(zg_ctx*, const zg_fr*, zg_fr*, uint32_t, size_t, size_t);
extern "C" int zg_coeff_to_extended_dev(zg_ctx*, const zg_fr*, size_t, uint32_t, uint32_t, zg_fr*, size_t, size_t);
extern "C" int zg_extended_to_coeff_dev(zg_ctx*, const zg_fr*, uint32_t, uint32_t, size_t, zg_fr*);
namespace zg {
int ntt_halo_coset_dev(zg_ctx* ctx, const Fr* coeff, uint32_t n_in, uint32_t ext_k, Fr* out);
}

namespace {

// ---- keccak256 ----------------------------------------------------------------------------------
static const uint64_t KRC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static inline uint64_t rotl64(uint64_t v, int r) { return r ? (v << r) | (v >> (64 - r)) : v; }

static void keccak_f(uint64_t a[25]) {  // a[x + 5*y]
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(a[x + 5 * y], KROT[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= KRC[round];
  }
}
static void keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
  const size_t rate = 136;
  uint64_t a[25] = {0};
  std::vector<uint8_t> p(data, data + len);
  p.push_back(0x01);
  while (p.size() % rate) p.push_back(0);
  p.back() |= 0x80;
  for (size_t off = 0; off < p.size(); off += rate) {
    for (size_t i = 0; i < rate / 8; i++) {
      uint64_t w;
      memcpy(&w, &p[off + 8 * i], 8);
      a[i] ^= w;
    }
    keccak_f(a);
  }
  memcpy(out, a, 32);
}

// ---- host field helpers ---------------------------------------------------------------------------
static Fr fr_one() { return fp_one<FrParams>(); }
static Fr fr_u64(uint64_t x) { return fp_from_u64<FrParams>(x); }
static Fr fr_delta() {
  Fr raw;
  const uint32_t v[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
  for (int i = 0; i < 8; i++) raw.v[i] = v[i];
  return fp_to_mont(raw);
}
template <class P>
static void to_be32(const Fp<P>& mont, uint8_t out[32]) {
  Fp<P> c = fp_from_mont(mont);
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) out[31 - (4 * i + b)] = (uint8_t)(c.v[i] >> (8 * b));
}
static Fr fr_from_be32_mod(const uint8_t h[32]) {
  // int(h, big endian) mod r -> Montgomery; the raw value can exceed r, so it is the second operand
  Fr raw;
  for (int i = 0; i < 8; i++) {
    uint32_t w = 0;
    for (int b = 0; b < 4; b++) w |= (uint32_t)h[31 - (4 * i + b)] << (8 * b);
    raw.v[i] = w;
  }
  Fr r2;
  for (int i = 0; i < 8; i++) r2.v[i] = FrParams::r2(i);
  return fp_mul(r2, raw);
}
static Fr fr_pow(const Fr& a, uint64_t e) { return fp_pow_var(a, e); }

struct Affine {
  Fq x, y;
  bool inf;
};
// Curve::batch_normalize for a handful of points
static void batch_normalize(const G1Jac* in, size_t n, Affine* out) {
  std::vector<Fq> pre(n);
  Fq acc = fp_one<FqParams>();
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    if (!fp_is_zero(in[i].z)) acc = fp_mul(acc, in[i].z);
  }
  Fq inv = fp_inv(acc);
  for (size_t i = n; i-- > 0;) {
    if (fp_is_zero(in[i].z)) {
      out[i].inf = true;
      out[i].x = out[i].y = fp_zero<FqParams>();
      continue;
    }
    Fq zi = fp_mul(inv, pre[i]);
    inv = fp_mul(inv, in[i].z);
    Fq zi2 = fp_sqr(zi);
    out[i].x = fp_mul(in[i].x, zi2);
    out[i].y = fp_mul(in[i].y, fp_mul(zi2, zi));
    out[i].inf = false;
  }
}

// snark_verifier EvmTranscript<G1Affine, NativeLoader, _, Vec<u8>>
struct Transcript {
  std::vector<uint8_t> buf, out;
  void common_scalar(const Fr& s) {
    uint8_t b[32];
    to_be32(s, b);
    buf.insert(buf.end(), b, b + 32);
  }
  void write_scalar(const Fr& s) {
    uint8_t b[32];
    to_be32(s, b);
    buf.insert(buf.end(), b, b + 32);
    out.insert(out.end(), b, b + 32);
  }
  bool write_point(const Affine& p) {
    if (p.inf) return false;  // "Cannot write points at infinity to the transcript"
    uint8_t b[64];
    to_be32(p.x, b);
    to_be32(p.y, b + 32);
    buf.insert(buf.end(), b, b + 64);
    out.insert(out.end(), b, b + 64);
    return true;
  }
  Fr squeeze() {
    std::vector<uint8_t> d(buf);
    if (buf.size() == 32) d.push_back(1);
    uint8_t h[32];
    keccak256(d.data(), d.size(), h);
    buf.assign(h, h + 32);
    return fr_from_be32_mod(h);
  }
};

struct Bump {
  uint8_t* base = nullptr;
  size_t off = 0, cap = 0;
  template <class T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T* p = (T*)(base ? base + off : nullptr);
    off += bytes;
    return p;
  }
};

}  // namespace

// The read-only half of a proving key (fixed / sigma columns in three forms, l_0 / l_last / l_active / X, domain tables,
// programs): shared by the clones of a key (zg_pk_clone), freed with the last of them.
struct PkResident {
  std::atomic<int> refs{1};
  uint8_t* arena = nullptr;
  ~PkResident() { if (arena) cudaFree(arena); }
};

struct zg_pk {
  uint32_t k = 0, ext_k = 0, rot_scale = 0;   // ext_k: halo2's extended_k (only the C-ABI conversions of zg_evaluate_h use it)
  size_t n = 0, N = 0;                        // N = ext.N rows of the internal extended domain
  ExtDomain ext;
  Fr* ext_mem = nullptr;
  uint32_t blk_first = 0, blk_step = 1;       // coset blocks this rank computes in the current proof (all: 0, 1)
  uint32_t A = 0, F = 0, I = 0, degree = 0, bf = 0, usable = 0, qdeg = 0;
  std::vector<std::pair<uint32_t, int32_t>> q[3];
  std::vector<std::pair<uint32_t, uint32_t>> perm;  // (kind, index)
  uint32_t m = 0, chunk = 0, nsets = 0;
  uint32_t n_progs = 0, n_gate_progs = 0, n_lookups = 0;
  std::vector<uint32_t> in_first, in_count, tab_first, tab_count;
  Fr transcript_repr, omega, omega_inv, delta;
  std::vector<G1Affine> fixed_comm, sigma_comm;
  PkResident* res = nullptr;    // resident half (shared)
  uint8_t* arena = nullptr;     // per-proof workspace half (this key's own)
  size_t n_ops = 0, n_prog_off = 0, n_constants = 0, n_queries = 0;
  // resident
  Fr *fixed_values, *fixed_polys, *fixed_cosets, *sigma_values, *sigma_polys, *sigma_cosets;
  Fr *l0, *l_last, *l_active, *coset_x, *t_inv, *constants, *omega_pows;
  uint32_t *ops, *prog_off, *d_in_first, *d_in_count, *d_tab_first, *d_tab_count;
  uint32_t* d_qcol[3];
  int32_t* d_qrot[3];
  const Fr** d_cols_base[3];
  const Fr** d_cols_ext[3];
  const Fr** d_perm_cosets;   // m
  const Fr** d_sigma_cosets;  // m
  const Fr** d_z_cosets;      // nsets
  // per-proof workspace
  Fr *adv_values, *adv_polys, *adv_cosets, *inst_values, *inst_polys, *inst_cosets;
  Fr *ci, *ct, *pa, *ps, *pa_poly, *ps_poly, *lz, *lz_poly, *lk_cosets;
  Fr *pz, *pz_poly, *pz_coset, *frac, *rnd, *random_poly, *h, *h_coeff, *h_poly, *fold, *wpoly, *scratch, *evals_dev, *points_dev,
      *coeff_dev, *one_dev;
  const Fr** d_polyptrs;
  uint32_t* d_pidx;
  uint32_t* d_status;
  uint8_t* lookup_ws;
  size_t lookup_ws_stride = 0;
  uint8_t* table_cache;
  size_t table_cache_stride = 0;
  std::vector<char> table_cached;
  uint64_t* rnd_words_dev;
  uint64_t* rnd_words_host = nullptr;  // pinned
  size_t n_draws = 0;
  float stage_ms[8] = {0};
  std::vector<char> table_cacheable;   // per lookup: the table reads fixed columns and constants only
  zg_pk() = default;
  zg_pk(const zg_pk&) = default;       // only zg_pk_clone copies, and it re-points every owned resource at once
  zg_pk& operator=(const zg_pk&) = delete;
  ~zg_pk() {                           // every early return of zg_pk_load / zg_pk_clone releases through this
    if (res && res->refs.fetch_sub(1) == 1) delete res;
    if (arena) cudaFree(arena);
    if (rnd_words_host) cudaFreeHost(rnd_words_host);
  }
};

namespace {

static int parse_cs(zg_ctx* ctx, zg_pk* pk, const zg_pk_desc* d, std::vector<uint32_t>& ops, std::vector<uint32_t>& prog_off) {
  const uint32_t* w = d->cs_words;
  size_t nw = d->cs_nwords, p = 0;
  auto need = [&](size_t c) { return p + c <= nw; };
  if (!need(6) || w[0] != 0x5A473031u) return ctx->fail(ZG_E_INVALID, "pk_load: bad constraint-system blob");
  pk->A = w[1]; pk->F = w[2]; pk->I = w[3]; pk->degree = w[4]; pk->bf = w[5];
  p = 6;
  for (int kind = 0; kind < 3; kind++) {
    if (!need(1)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
    uint32_t c = w[p++];
    if (!need(2 * (size_t)c)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
    for (uint32_t i = 0; i < c; i++, p += 2) pk->q[kind].push_back({w[p], (int32_t)w[p + 1]});
  }
  if (!need(1)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
  pk->m = w[p++];
  if (!need(2 * (size_t)pk->m)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
  for (uint32_t i = 0; i < pk->m; i++, p += 2) pk->perm.push_back({w[p], w[p + 1]});
  if (!need(1)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
  pk->n_progs = w[p++];
  if (!need(pk->n_progs + 3)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
  prog_off.assign(w + p, w + p + pk->n_progs + 1);
  p += pk->n_progs + 1;
  pk->n_gate_progs = w[p++];
  pk->n_lookups = w[p++];
  if (!need(4 * (size_t)pk->n_lookups + 1)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
  for (uint32_t l = 0; l < pk->n_lookups; l++, p += 4) {
    pk->in_first.push_back(w[p]); pk->in_count.push_back(w[p + 1]);
    pk->tab_first.push_back(w[p + 2]); pk->tab_count.push_back(w[p + 3]);
  }
  uint32_t nops = w[p++];
  if (!need(nops)) return ctx->fail(ZG_E_INVALID, "pk_load: truncated blob");
  ops.assign(w + p, w + p + nops);
  if (pk->degree < 3) return ctx->fail(ZG_E_INVALID, "pk_load: constraint degree < 3");
  // every program: operands in range, stack depth within the interpreter's (expr.cuh EXPR_STACK), one value left
  for (uint32_t g = 0; g < pk->n_progs; g++) {
    if (prog_off[g] > prog_off[g + 1] || prog_off[g + 1] > nops) return ctx->fail(ZG_E_INVALID, "pk_load: bad program table");
    int depth = 0;
    for (uint32_t pc = prog_off[g]; pc < prog_off[g + 1]; pc++) {
      const uint32_t op = ops[pc] & 0xff, arg = ops[pc] >> 8;
      switch (op) {
        case OP_CONST: if (arg >= d->n_constants) return ctx->fail(ZG_E_INVALID, "pk_load: constant index out of range"); depth++; break;
        case OP_ADVICE: case OP_FIXED: case OP_INSTANCE:
          if (arg >= pk->q[op - OP_ADVICE].size()) return ctx->fail(ZG_E_INVALID, "pk_load: query index out of range");
          depth++; break;
        case OP_NEG: if (depth < 1) return ctx->fail(ZG_E_INVALID, "pk_load: malformed program"); break;
        case OP_SCALE:
          if (depth < 1 || arg >= d->n_constants) return ctx->fail(ZG_E_INVALID, "pk_load: malformed program");
          break;
        case OP_ADD: case OP_SUB: case OP_MUL:
          if (depth < 2) return ctx->fail(ZG_E_INVALID, "pk_load: malformed program");
          depth--; break;
        default: return ctx->fail(ZG_E_INVALID, "pk_load: unknown opcode");
      }
      if (depth > EXPR_STACK) return ctx->fail(ZG_E_INVALID, "pk_load: expression deeper than the evaluator's stack");
    }
    if (depth != 1) return ctx->fail(ZG_E_INVALID, "pk_load: malformed program");
  }
  const uint32_t ncols[3] = {pk->A, pk->F, pk->I};
  for (int kind = 0; kind < 3; kind++)
    for (auto& qq : pk->q[kind])
      if (qq.first >= ncols[kind]) return ctx->fail(ZG_E_INVALID, "pk_load: query names a column that does not exist");
  // a lookup table may be sorted once per key only if nothing in it changes from proof to proof
  pk->table_cacheable.assign(pk->n_lookups, 0);
  for (uint32_t l = 0; l < pk->n_lookups; l++) {
    if ((uint64_t)pk->in_first[l] + pk->in_count[l] > pk->n_progs || (uint64_t)pk->tab_first[l] + pk->tab_count[l] > pk->n_progs ||
        !pk->in_count[l] || !pk->tab_count[l])
      return ctx->fail(ZG_E_INVALID, "pk_load: lookup program range out of bounds");
    bool fixed_only = pk->tab_count[l] == 1;
    for (uint32_t g = pk->tab_first[l]; fixed_only && g < pk->tab_first[l] + pk->tab_count[l]; g++)
      for (uint32_t pc = prog_off[g]; pc < prog_off[g + 1]; pc++) {
        const uint32_t op = ops[pc] & 0xff;
        if (op == OP_ADVICE || op == OP_INSTANCE) fixed_only = false;
      }
    pk->table_cacheable[l] = fixed_only;
  }
  return ZG_OK;
}

template <class T>
static cudaError_t upload(T* dst, const std::vector<T>& v, cudaStream_t st) {
  if (v.empty()) return cudaSuccess;
  return cudaMemcpyAsync(dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

// commit `count` columns (stride n) on `basis`, normalise on the host
static int commit_batch(zg_ctx* ctx, int basis, const Fr* cols, size_t stride, size_t n, size_t count, Affine* out,
                        bool round_of_a_proof = false) {
  // keygen commits locally on every rank; the rounds of a proof may be spread over the ranks (dist.cu)
  int rc = round_of_a_proof ? msm_round(ctx, basis, cols, stride, n, count, 0)
                            : zg_msm_dev(ctx, basis, (const zg_fr*)cols, stride, n, count, (zg_g1*)ctx->d_msm_out);
  if (rc) return rc;
  std::vector<G1Jac> jac(count);
  ZG_CUDA(cudaMemcpyAsync(jac.data(), ctx->d_msm_out, sizeof(G1Jac) * count, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  batch_normalize(jac.data(), count, out);
  return ZG_OK;
}


// Work enqueued while an AuxScope is alive goes to the context's low-priority stream; it starts after everything
// enqueued on the proof's main stream so far (fork) and is awaited by join_aux before the quotient stage.
struct AuxScope {
  zg_ctx* ctx;
  cudaStream_t saved;
  cudaError_t err;
  explicit AuxScope(zg_ctx* c) : ctx(c), saved(c->stream) {
    err = cudaEventRecord(c->ev_fork, c->stream);
    if (err == cudaSuccess) err = cudaStreamWaitEvent(c->aux, c->ev_fork, 0);
    c->stream = c->aux;
  }
  ~AuxScope() { ctx->stream = saved; }
};
static cudaError_t join_aux(zg_ctx* ctx) {
  cudaError_t e = cudaEventRecord(ctx->ev_join, ctx->aux);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
  return e;
}

// environment of the RPN programs over the 2^k domain (ext = false) or the extended coset (ext = true)
static ExprEnv make_env(const zg_pk* pk, bool ext) {
  ExprEnv e;
  for (int kind = 0; kind < 3; kind++) {
    e.cols[kind] = ext ? pk->d_cols_ext[kind] : pk->d_cols_base[kind];
    e.qcol[kind] = pk->d_qcol[kind];
    e.qrot[kind] = pk->d_qrot[kind];
  }
  e.constants = pk->constants;
  e.ops = pk->ops;
  e.size = (uint32_t)(ext ? pk->N : pk->n);
  e.rot_scale = ext ? pk->rot_scale : 1;
  e.wrap_mask = (uint32_t)(ext ? pk->ext.B - 1 : pk->n - 1);
  return e;
}

// extended-coset forms of the per-proof polynomials (coefficient form -> zeta * <omega_ext>), in the batches the prover
// builds them: advice + instance, lookup permuted [a | s], product polynomials (permutation z, lookup z)
static int coset_advice_instance(zg_ctx* ctx, zg_pk* pk) {
  int rc = ext_from_coeff(ctx, pk->ext, pk->adv_polys, pk->n, pk->adv_cosets, pk->N, pk->A, pk->blk_first, pk->blk_step);
  if (rc || !pk->I) return rc;
  return ext_from_coeff(ctx, pk->ext, pk->inst_polys, pk->n, pk->inst_cosets, pk->N, pk->I, pk->blk_first, pk->blk_step);
}
static int coset_lookup_permuted(zg_ctx* ctx, zg_pk* pk) {
  if (!pk->n_lookups) return ZG_OK;
  return ext_from_coeff(ctx, pk->ext, pk->pa_poly, pk->n, pk->lk_cosets, pk->N, 2 * pk->n_lookups, pk->blk_first, pk->blk_step);
}
static int coset_products(zg_ctx* ctx, zg_pk* pk) {
  int rc = ZG_OK;
  if (pk->nsets)
    rc = ext_from_coeff(ctx, pk->ext, pk->pz_poly, pk->n, pk->pz_coset, pk->N, pk->nsets, pk->blk_first, pk->blk_step);
  if (rc || !pk->n_lookups) return rc;
  return ext_from_coeff(ctx, pk->ext, pk->lz_poly, pk->n, pk->lk_cosets + 2 * (size_t)pk->n_lookups * pk->N, pk->N, pk->n_lookups,
                        pk->blk_first, pk->blk_step);
}

// Evaluator::evaluate_h: reads the coefficient forms in the pk workspace (adv_polys, inst_polys, pz_poly, pa_poly | ps_poly,
// lz_poly), builds their extended cosets and leaves the y-folded numerator in pk->h.
// `transform` = false when the caller has already built the cosets (create_proof does so early, off the critical path).
static int quotient_numerator(zg_ctx* ctx, zg_pk* pk, const Fr& theta, const Fr& beta, const Fr& gamma, const Fr& y,
                              bool transform) {
  const size_t n = pk->n, N = pk->N;
  const uint32_t A = pk->A, I = pk->I, Lk = pk->n_lookups, S = pk->nsets, bf = pk->bf, m = pk->m;
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  LookupProgs lp{pk->prog_off, pk->d_in_first, pk->d_in_count, pk->d_tab_first, pk->d_tab_count};
  int rc;
  if (transform) {
    rc = coset_advice_instance(ctx, pk);
    if (rc) return rc;
    rc = coset_products(ctx, pk);
    if (rc) return rc;
    rc = coset_lookup_permuted(ctx, pk);
    if (rc) return rc;
  }
  // rows: the whole domain, or -- when the proof is spread over GPUs -- the coset blocks this rank owns (rotations
  // never leave a block, so a block's rows need nothing from the others)
  const bool all_rows = pk->blk_first == 0 && pk->blk_step == 1;
  for (uint32_t c = pk->blk_first; c < (all_rows ? 1u : pk->ext.cosets); c += pk->blk_step) {
    ExprEnv ext_env = make_env(pk, true);
    if (!all_rows) {
      ext_env.row0 = (uint32_t)(c * pk->ext.B);
      ext_env.size = (uint32_t)((c + 1) * pk->ext.B);
    }
    expr_h_gates(ext_env, pk->prog_off, pk->n_gate_progs, y, pk->h, st, lc);
    if (S) {
      PermEnv pe;
      pe.z_cosets = pk->d_z_cosets; pe.col_cosets = pk->d_perm_cosets; pe.sigma_cosets = pk->d_sigma_cosets;
      pe.l0 = pk->l0; pe.l_last = pk->l_last; pe.l_active = pk->l_active; pe.coset_x = pk->coset_x;
      pe.nsets = S; pe.m = m; pe.chunk = pk->chunk; pe.rot_scale = pk->rot_scale;
      pe.row0 = ext_env.row0; pe.size = ext_env.size;
      pe.wrap_mask = (uint32_t)(pk->ext.B - 1);
      pe.last_rot = -(int32_t)(bf + 1);
      expr_h_permutation(pe, beta, gamma, y, pk->delta, pk->h, st, lc);
    }
    if (Lk) {
      Fr* as_cos = pk->lk_cosets;                           // [a | s] (2 Lk columns), then z (Lk)
      Fr* z_cos = pk->lk_cosets + 2 * (size_t)Lk * N;
      for (uint32_t l = 0; l < Lk; l++) {
        LookupHEnv le{z_cos + l * N, as_cos + l * N, as_cos + (Lk + l) * N, pk->l0, pk->l_last, pk->l_active};
        expr_h_lookup(ext_env, lp, l, le, theta, beta, gamma, y, pk->h, st, lc);
      }
    }
  }
  return ZG_OK;
}


// ---- device memory of a key: resident half (shared by clones) and workspace half (one per key) ---------------------------
static size_t carve_resident(zg_pk* pk, uint8_t* base) {
  const size_t n = pk->n, N = pk->N;
  const uint32_t F = pk->F, m = pk->m, Lk = pk->n_lookups;
  Bump b;
  b.base = base;
  pk->fixed_values = b.take<Fr>(F * n); pk->fixed_polys = b.take<Fr>(F * n); pk->fixed_cosets = b.take<Fr>(F * N);
  pk->sigma_values = b.take<Fr>(m * n); pk->sigma_polys = b.take<Fr>(m * n); pk->sigma_cosets = b.take<Fr>(m * N);
  pk->l0 = b.take<Fr>(N); pk->l_last = b.take<Fr>(N); pk->l_active = b.take<Fr>(N); pk->coset_x = b.take<Fr>(N);
  pk->t_inv = b.take<Fr>((size_t)1 << (pk->ext_k - pk->k)); pk->ext_mem = b.take<Fr>(ext_domain_table_elems(pk->ext));
  pk->constants = b.take<Fr>(pk->n_constants + 1); pk->omega_pows = b.take<Fr>(n);
  pk->ops = b.take<uint32_t>(pk->n_ops + 1); pk->prog_off = b.take<uint32_t>(pk->n_prog_off);
  pk->d_in_first = b.take<uint32_t>(Lk + 1); pk->d_in_count = b.take<uint32_t>(Lk + 1);
  pk->d_tab_first = b.take<uint32_t>(Lk + 1); pk->d_tab_count = b.take<uint32_t>(Lk + 1);
  for (int kind = 0; kind < 3; kind++) {
    pk->d_qcol[kind] = b.take<uint32_t>(pk->q[kind].size() + 1);
    pk->d_qrot[kind] = b.take<int32_t>(pk->q[kind].size() + 1);
  }
  pk->d_sigma_cosets = b.take<const Fr*>(m + 1);
  return b.off;
}
static size_t carve_workspace(zg_pk* pk, uint8_t* base) {
  const size_t n = pk->n, N = pk->N;
  const uint32_t A = pk->A, F = pk->F, I = pk->I, m = pk->m, Lk = pk->n_lookups, S = pk->nsets;
  const size_t n_queries = pk->n_queries;
  Bump b;
  b.base = base;
  const uint32_t ncols[3] = {A, F, I};
  for (int kind = 0; kind < 3; kind++) {
    pk->d_cols_base[kind] = b.take<const Fr*>(ncols[kind] + 1);
    pk->d_cols_ext[kind] = b.take<const Fr*>(ncols[kind] + 1);
  }
  pk->d_perm_cosets = b.take<const Fr*>(m + 1);
  pk->d_z_cosets = b.take<const Fr*>(S + 1);
  pk->adv_values = b.take<Fr>(A * n); pk->adv_polys = b.take<Fr>(A * n); pk->adv_cosets = b.take<Fr>(A * N);
  pk->inst_values = b.take<Fr>(I * n); pk->inst_polys = b.take<Fr>(I * n); pk->inst_cosets = b.take<Fr>(I * N);
  pk->ci = b.take<Fr>(Lk * n); pk->ct = b.take<Fr>(Lk * n);
  pk->pa = b.take<Fr>(2 * Lk * n); pk->ps = pk->pa ? pk->pa + (size_t)Lk * n : nullptr;   // pa | ps contiguous: one MSM batch
  pk->pa_poly = b.take<Fr>(2 * Lk * n); pk->ps_poly = pk->pa_poly ? pk->pa_poly + (size_t)Lk * n : nullptr;
  pk->pz = b.take<Fr>((S + Lk + 1) * n); pk->lz = pk->pz ? pk->pz + (size_t)S * n : nullptr;   // pz | lz | (random poly) contiguous
  pk->pz_poly = b.take<Fr>((S + Lk) * n); pk->lz_poly = pk->pz_poly ? pk->pz_poly + (size_t)S * n : nullptr;
  pk->pz_coset = b.take<Fr>(S * N); pk->lk_cosets = b.take<Fr>(3 * (size_t)Lk * N);
  pk->frac = b.take<Fr>(n); pk->rnd = b.take<Fr>(pk->n_draws); pk->random_poly = pk->rnd;  // set per proof
  const size_t hcap = std::max<size_t>(N, (size_t)(S + Lk + 1) * n);   // also scratch for the S + Lk fraction columns
  pk->h = b.take<Fr>(hcap); pk->h_coeff = b.take<Fr>(hcap); pk->h_poly = b.take<Fr>(n);
  pk->fold = b.take<Fr>(8 * n); pk->wpoly = b.take<Fr>(8 * n);
  pk->scratch = b.take<Fr>(8 * 4096 + std::max<size_t>(64, (n + 4095) / 4096) * n_queries);
  pk->evals_dev = b.take<Fr>(n_queries + 8); pk->points_dev = b.take<Fr>(16); pk->coeff_dev = b.take<Fr>(n_queries + 8);
  pk->one_dev = b.take<Fr>(8);
  pk->d_polyptrs = b.take<const Fr*>(std::max<size_t>(n_queries + 8, 2 * (size_t)m + 8)); pk->d_pidx = b.take<uint32_t>(n_queries + 8);
  pk->d_status = b.take<uint32_t>(64);
  pk->lookup_ws_stride = (lookup_workspace_bytes(pk->usable) + 255) & ~(size_t)255;
  pk->lookup_ws = b.take<uint8_t>(pk->lookup_ws_stride * zg_ctx::N_SIDE);
  pk->table_cache_stride = (lookup_table_bytes(pk->usable) + 255) & ~(size_t)255;
  pk->table_cache = b.take<uint8_t>(pk->table_cache_stride * (Lk + 1));     // per key: filled lazily by its first proof
  pk->rnd_words_dev = b.take<uint64_t>(8 * pk->n_draws);
  return b.off;
}
// the pointer tables that name workspace columns (and the constant 1) -- once per key, also for a clone
static int upload_workspace_tables(zg_ctx* ctx, zg_pk* pk) {
  const size_t n = pk->n, N = pk->N;
  const uint32_t A = pk->A, F = pk->F, I = pk->I, m = pk->m, S = pk->nsets;
  cudaStream_t st = ctx->stream;
  std::vector<const Fr*> pb[3], pe[3];
  for (uint32_t c = 0; c < A; c++) { pb[0].push_back(pk->adv_values + c * n); pe[0].push_back(pk->adv_cosets + c * N); }
  for (uint32_t c = 0; c < F; c++) { pb[1].push_back(pk->fixed_values + c * n); pe[1].push_back(pk->fixed_cosets + c * N); }
  for (uint32_t c = 0; c < I; c++) { pb[2].push_back(pk->inst_values + c * n); pe[2].push_back(pk->inst_cosets + c * N); }
  for (int kind = 0; kind < 3; kind++) {
    ZG_CUDA(upload(pk->d_cols_base[kind], pb[kind], st));
    ZG_CUDA(upload(pk->d_cols_ext[kind], pe[kind], st));
  }
  std::vector<const Fr*> pcos, zcos;
  for (uint32_t c = 0; c < m; c++) {
    uint32_t kind = pk->perm[c].first, idx = pk->perm[c].second;
    if (kind > 2 || idx >= (kind == 0 ? A : kind == 1 ? F : I)) return ctx->fail(ZG_E_INVALID, "pk_load: bad permutation column");
    pcos.push_back(pe[kind][idx]);
  }
  for (uint32_t s = 0; s < S; s++) zcos.push_back(pk->pz_coset + s * N);
  ZG_CUDA(upload(pk->d_perm_cosets, pcos, st));
  ZG_CUDA(upload(pk->d_z_cosets, zcos, st));
  Fr onev = fr_one();
  ZG_CUDA(cudaMemcpyAsync(pk->one_dev, &onev, sizeof(Fr), cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaStreamSynchronize(st));          // the vectors above are host temporaries
  return ZG_OK;
}

}  // namespace

extern "C" {

void zg_xorshift_seed(zg_xorshift* r, const uint8_t seed[16]) {
  uint32_t s[4];
  memcpy(s, seed, 16);
  if ((s[0] | s[1] | s[2] | s[3]) == 0) s[0] = s[1] = s[2] = s[3] = 0x0BAD5EEDu;
  r->x = s[0]; r->y = s[1]; r->z = s[2]; r->w = s[3];
}
void zg_xorshift_fill(void* state, uint64_t* out, size_t n) {
  zg_xorshift* r = (zg_xorshift*)state;
  uint32_t x = r->x, y = r->y, z = r->z, w = r->w;
  for (size_t i = 0; i < n; i++) {
    uint32_t t = x ^ (x << 11);
    x = y; y = z; z = w;
    w = w ^ (w >> 19) ^ (t ^ (t >> 8));
    uint64_t lo = w;
    t = x ^ (x << 11);
    x = y; y = z; z = w;
    w = w ^ (w >> 19) ^ (t ^ (t >> 8));
    out[i] = lo | ((uint64_t)w << 32);
  }
  r->x = x; r->y = y; r->z = z; r->w = w;
}

void zg_pk_free(zg_ctx* ctx, zg_pk* pk) {
  if (!pk) return;
  cudaStreamSynchronize(ctx->stream);
  delete pk;
}

int zg_pk_last_stage_ms(const zg_pk* pk, float out[8]) {
  if (!pk) return ZG_E_INVALID;
  memcpy(out, pk->stage_ms, sizeof(float) * 8);
  return ZG_OK;
}

// the transcript's hash (host code), exposed so that CPU-only tests can check it against the published keccak256 vectors
void zg_debug_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) { keccak256(data, len, out); }

int zg_pk_set_transcript_repr(zg_ctx* ctx, zg_pk* pk, const zg_fr* transcript_repr) {
  ZG_ENTER(ctx);
  if (!pk || !transcript_repr) return ctx->fail(ZG_E_INVALID, "pk_set_transcript_repr: null argument");
  memcpy(pk->transcript_repr.v, transcript_repr, 32);
  return ZG_OK;
}

int zg_pk_commitments(zg_ctx* ctx, const zg_pk* pk, zg_g1_affine* fixed_out, zg_g1_affine* sigma_out) {
  ZG_ENTER(ctx);
  if (!pk) return ctx->fail(ZG_E_INVALID, "pk_commitments: null pk");
  if (fixed_out) memcpy(fixed_out, pk->fixed_comm.data(), sizeof(G1Affine) * pk->F);
  if (sigma_out) memcpy(sigma_out, pk->sigma_comm.data(), sizeof(G1Affine) * pk->m);
  return ZG_OK;
}

// A second key over the same resident columns: own workspace, own lookup-table cache, own pinned RNG staging.  The clone
// may live on another context of the SAME device (one lane of a ProofService each); the resident half is freed with the
// last key that uses it.
int zg_pk_clone(zg_ctx* ctx, const zg_pk* src, zg_pk** out) {
  ZG_ENTER(ctx);
  if (!src || !out) return ctx->fail(ZG_E_INVALID, "pk_clone: null argument");
  *out = nullptr;
  if (!ctx->srs_loaded || ctx->srs_k != src->k || !ctx->table[0].pts || !ctx->table[1].pts)
    return ctx->fail(ZG_E_STATE, "pk_clone: load (or share) an SRS with both bases for the same k first");
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, src->res->arena) != cudaSuccess || attr.device != ctx->device)
    return ctx->fail(ZG_E_INVALID, "pk_clone: the source key lives on another device");
  std::unique_ptr<zg_pk> pk(new zg_pk(*src));
  pk->arena = nullptr;                     // nothing of the source's own resources may be released by the copy
  pk->rnd_words_host = nullptr;
  pk->res->refs.fetch_add(1);
  pk->table_cached.assign(pk->n_lookups, 0);
  memset(pk->stage_ms, 0, sizeof(pk->stage_ms));
  const size_t bytes = carve_workspace(pk.get(), nullptr);
  if (cudaMalloc(&pk->arena, bytes) != cudaSuccess) return ctx->fail(ZG_E_NOMEM, "pk_clone: workspace allocation failed");
  carve_workspace(pk.get(), pk->arena);
  ZG_CUDA(cudaMallocHost(&pk->rnd_words_host, 8 * pk->n_draws * sizeof(uint64_t)));
  int rc = upload_workspace_tables(ctx, pk.get());
  if (rc) return rc;
  *out = pk.release();
  return ZG_OK;
}

int zg_pk_read_column(zg_ctx* ctx, const zg_pk* pk, int what, uint32_t index, zg_fr* out) {
  ZG_ENTER(ctx);
  if (!pk || !out) return ctx->fail(ZG_E_INVALID, "pk_read_column: null argument");
  const Fr* base;
  uint32_t count;
  switch (what) {
    case ZG_PK_FIXED_VALUES: base = pk->fixed_values; count = pk->F; break;
    case ZG_PK_FIXED_POLYS: base = pk->fixed_polys; count = pk->F; break;
    case ZG_PK_SIGMA_VALUES: base = pk->sigma_values; count = pk->m; break;
    case ZG_PK_SIGMA_POLYS: base = pk->sigma_polys; count = pk->m; break;
    default: return ctx->fail(ZG_E_INVALID, "pk_read_column: unknown column family");
  }
  if (index >= count) return ctx->fail(ZG_E_INVALID, "pk_read_column: index out of range");
  ZG_CUDA(cudaMemcpyAsync(out, base + (size_t)index * pk->n, sizeof(Fr) * pk->n, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}

int zg_pk_load(zg_ctx* ctx, const zg_pk_desc* d, zg_pk** out) {
  ZG_ENTER(ctx);
  if (!d || !out) return ctx->fail(ZG_E_INVALID, "pk_load: null argument");
  *out = nullptr;
  if (!ctx->srs_loaded || ctx->srs_k != d->k || !ctx->table[0].pts || !ctx->table[1].pts)
    return ctx->fail(ZG_E_STATE, "pk_load: load an SRS with both bases for the same k first");
  std::unique_ptr<zg_pk> pk(new zg_pk());
  std::vector<uint32_t> ops, prog_off;
  int rc = parse_cs(ctx, pk.get(), d, ops, prog_off);
  if (rc) return rc;
  pk->k = d->k;
  pk->n = (size_t)1 << d->k;
  pk->qdeg = pk->degree - 1;
  pk->ext_k = pk->k;
  while (((size_t)1 << pk->ext_k) < pk->n * pk->qdeg) pk->ext_k++;
  if (pk->ext_k > 28) return ctx->fail(ZG_E_INVALID, "pk_load: extended domain exceeds the field's 2-adicity");
  ext_domain_shape(pk->k, pk->qdeg, pk->ext);     // internal domain: 3 cosets of 2n for a degree-6 system (extdomain.cuh)
  pk->N = pk->ext.N;
  pk->rot_scale = pk->ext.rot_scale;
  pk->usable = (uint32_t)pk->n - (pk->bf + 1);
  pk->chunk = pk->degree - 2;
  pk->nsets = (pk->m + pk->chunk - 1) / pk->chunk;
  pk->omega = host_omega(pk->k);
  pk->omega_inv = fp_inv(pk->omega);
  pk->delta = fr_delta();
  memcpy(pk->transcript_repr.v, &d->transcript_repr, 32);
  pk->table_cached.assign(pk->n_lookups, 0);
  if (pk->n < pk->bf + 3) return ctx->fail(ZG_E_INVALID, "pk_load: not enough rows");
  const size_t n = pk->n, N = pk->N;
  const uint32_t A = pk->A, F = pk->F, I = pk->I, m = pk->m, Lk = pk->n_lookups, S = pk->nsets;
  if (A + I + F > 60 || m > 60) return ctx->fail(ZG_E_INVALID, "pk_load: too many columns for one commitment batch");
  // RNG draws of one proof, in order
  pk->n_draws = (size_t)A * (pk->bf + 1) + A + (size_t)Lk * (2 * (pk->bf + 1) + 2) + (size_t)S * (pk->bf + 1) +
                (size_t)Lk * (pk->bf + 1) + (n + 1) + pk->qdeg;
  const uint32_t n_queries = (uint32_t)(pk->q[0].size() + pk->q[1].size() + m + 3 * S + 5 * Lk + 2);

  pk->n_ops = ops.size(); pk->n_prog_off = prog_off.size(); pk->n_constants = d->n_constants; pk->n_queries = n_queries;
  // two arenas, each sized by a dry pass and carved by a second one: the resident half and the per-proof workspace
  pk->res = new PkResident();
  {
    size_t bytes = carve_resident(pk.get(), nullptr);
    cudaError_t e = cudaMalloc(&pk->res->arena, bytes);
    if (e != cudaSuccess) return ctx->fail(ZG_E_NOMEM, "pk_load: device arena allocation failed (resident half)");
    carve_resident(pk.get(), pk->res->arena);
    bytes = carve_workspace(pk.get(), nullptr);
    e = cudaMalloc(&pk->arena, bytes);
    if (e != cudaSuccess) return ctx->fail(ZG_E_NOMEM, "pk_load: device arena allocation failed (workspace half)");
    carve_workspace(pk.get(), pk->arena);
  }
  ZG_CUDA(cudaMallocHost(&pk->rnd_words_host, 8 * pk->n_draws * sizeof(uint64_t)));
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};

  // constraint system tables
  ZG_CUDA(upload(pk->ops, ops, st));
  ZG_CUDA(upload(pk->prog_off, prog_off, st));
  ZG_CUDA(upload(pk->d_in_first, pk->in_first, st)); ZG_CUDA(upload(pk->d_in_count, pk->in_count, st));
  ZG_CUDA(upload(pk->d_tab_first, pk->tab_first, st)); ZG_CUDA(upload(pk->d_tab_count, pk->tab_count, st));
  if (d->n_constants)
    ZG_CUDA(cudaMemcpyAsync(pk->constants, d->constants, sizeof(Fr) * d->n_constants, cudaMemcpyHostToDevice, st));
  std::vector<uint32_t> qc[3];
  std::vector<int32_t> qr[3];
  for (int kind = 0; kind < 3; kind++) {
    for (auto& pr : pk->q[kind]) { qc[kind].push_back(pr.first); qr[kind].push_back(pr.second); }
    ZG_CUDA(upload(pk->d_qcol[kind], qc[kind], st));
    ZG_CUDA(upload(pk->d_qrot[kind], qr[kind], st));
  }
  // pointer tables: sigma cosets live in the resident half, the others name workspace columns
  {
    std::vector<const Fr*> scos;
    for (uint32_t c = 0; c < m; c++) scos.push_back(pk->sigma_cosets + c * N);
    ZG_CUDA(upload(pk->d_sigma_cosets, scos, st));
    ZG_CUDA(cudaStreamSynchronize(st));
  }
  rc = upload_workspace_tables(ctx, pk.get());
  if (rc) return rc;
  Fr onev = fr_one();

  rc = ext_domain_init(ctx, pk->ext, pk->ext_mem, pk->scratch, pk->h);
  if (rc) return rc;
  // fixed columns: values -> coeff -> extended; commitments (keygen_vk)
  for (uint32_t c = 0; c < F; c++)
    ZG_CUDA(cudaMemcpyAsync(pk->fixed_values + c * n, d->fixed[c], sizeof(Fr) * n, cudaMemcpyHostToDevice, st));
  if (F) {
    rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)pk->fixed_values, (zg_fr*)pk->fixed_polys, pk->k, F, n);
    if (rc) return rc;
    rc = ext_from_coeff(ctx, pk->ext, pk->fixed_polys, n, pk->fixed_cosets, N, F);
    if (rc) return rc;
    std::vector<Affine> aff(F);
    rc = commit_batch(ctx, ZG_BASIS_LAGRANGE, pk->fixed_values, n, n, F, aff.data());
    if (rc) return rc;
    pk->fixed_comm.resize(F);
    for (uint32_t c = 0; c < F; c++) { pk->fixed_comm[c].x = aff[c].x; pk->fixed_comm[c].y = aff[c].y; }
  }
  // permutation: sigma values from the mapping
  if (m) {
    if (d->perm_mapping) {
      uint32_t* map_dev = (uint32_t*)pk->h;  // scratch: 2*m*n words fit in N Fr when 8*m*n <= 32*N
      if ((size_t)8 * m * n > sizeof(Fr) * N) return ctx->fail(ZG_E_INVALID, "pk_load: permutation too wide for scratch");
      ZG_CUDA(cudaMemcpyAsync(map_dev, d->perm_mapping, (size_t)8 * m * n, cudaMemcpyHostToDevice, st));
      std::vector<Fr> dp(m);
      Fr cur = fr_one();
      for (uint32_t c = 0; c < m; c++) { dp[c] = cur; cur = fp_mul(cur, pk->delta); }
      ZG_CUDA(cudaMemcpyAsync(pk->coeff_dev, dp.data(), sizeof(Fr) * m, cudaMemcpyHostToDevice, st));
      ZG_CUDA(cudaStreamSynchronize(st));      // `dp` is a host temporary
      fr_sigma_values(map_dev, pk->coeff_dev, pk->omega, m, n, pk->sigma_values, st, lc);
    } else if (d->sigma_values) {              // a serialized key carries the sigma columns themselves
      for (uint32_t c = 0; c < m; c++) {
        if (!d->sigma_values[c]) return ctx->fail(ZG_E_INVALID, "pk_load: null sigma column");
        ZG_CUDA(cudaMemcpyAsync(pk->sigma_values + c * n, d->sigma_values[c], sizeof(Fr) * n, cudaMemcpyHostToDevice, st));
      }
    } else {
      return ctx->fail(ZG_E_INVALID, "pk_load: neither perm_mapping nor sigma_values given");
    }
    rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)pk->sigma_values, (zg_fr*)pk->sigma_polys, pk->k, m, n);
    if (rc) return rc;
    rc = ext_from_coeff(ctx, pk->ext, pk->sigma_polys, n, pk->sigma_cosets, N, m);
    if (rc) return rc;
    std::vector<Affine> aff(m);
    rc = commit_batch(ctx, ZG_BASIS_LAGRANGE, pk->sigma_values, n, n, m, aff.data());
    if (rc) return rc;
    pk->sigma_comm.resize(m);
    for (uint32_t c = 0; c < m; c++) { pk->sigma_comm[c].x = aff[c].x; pk->sigma_comm[c].y = aff[c].y; }
  }
  // l_0, l_last, l_blind -> l_active, on the extended coset (use pz / pz_poly / h as scratch)
  {
    Fr* lag = pk->wpoly;  // 3 columns of n
    Fr* coef = pk->wpoly + 3 * n;
    ZG_CUDA(cudaMemsetAsync(lag, 0, sizeof(Fr) * 3 * n, st));
    std::vector<uint32_t> rows;
    rows.push_back(0);
    rows.push_back((uint32_t)(n + (n - pk->bf - 1)));
    for (uint32_t r = 0; r < pk->bf; r++) rows.push_back((uint32_t)(2 * n + (n - pk->bf + r)));
    std::vector<Fr> ones(rows.size(), fr_one());
    ZG_CUDA(cudaMemcpyAsync(pk->d_pidx, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, st));
    ZG_CUDA(cudaMemcpyAsync(pk->evals_dev, ones.data(), ones.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));
    fr_scatter_rows(lag, pk->d_pidx, pk->evals_dev, (uint32_t)rows.size(), st, lc);
    rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)lag, (zg_fr*)coef, pk->k, 3, n);
    if (rc) return rc;
    rc = ext_from_coeff(ctx, pk->ext, coef, n, pk->l0, N, 1);
    if (rc) return rc;
    rc = ext_from_coeff(ctx, pk->ext, coef + n, n, pk->l_last, N, 1);
    if (rc) return rc;
    rc = ext_from_coeff(ctx, pk->ext, coef + 2 * n, n, pk->h, N, 1);
    if (rc) return rc;
    fr_one_minus_sum(pk->l_last, pk->h, pk->l_active, N, st, lc);
  }
  // coset points X_i = zeta * ext_omega^i : extended form of the polynomial X
  {
    Fr* coef = pk->wpoly;
    ZG_CUDA(cudaMemsetAsync(coef, 0, sizeof(Fr) * n, st));
    ZG_CUDA(cudaMemcpyAsync(coef + 1, &onev, sizeof(Fr), cudaMemcpyHostToDevice, st));
    rc = ext_from_coeff(ctx, pk->ext, coef, n, pk->coset_x, N, 1);
    if (rc) return rc;
  }
  // halo2's coset, for the C-ABI conversions only: t_inv[i] = (zeta * ext_omega^i)^n - 1 (NOT inverted), period 2^(ext_k - k)
  {
    std::vector<Fr> t((size_t)1 << (pk->ext_k - pk->k));
    Fr zn = fr_pow(host_fr_zeta(), n);
    Fr wn = fr_pow(host_omega(pk->ext_k), n);
    Fr cur = zn;
    for (size_t i = 0; i < t.size(); i++) {
      t[i] = fp_sub(cur, fr_one());
      cur = fp_mul(cur, wn);
    }
    ZG_CUDA(cudaMemcpyAsync(pk->t_inv, t.data(), sizeof(Fr) * t.size(), cudaMemcpyHostToDevice, st));
  }
  // omega^i column for the permutation numerators
  fr_fill(pk->frac, pk->omega, n, st, lc);
  fr_running_product(pk->frac, pk->one_dev, pk->omega_pows, n, pk->scratch, st, lc);
  ZG_CUDA(cudaStreamSynchronize(st));
  *out = pk.release();
  return ZG_OK;
}

static int create_proof_impl(zg_ctx* ctx, zg_pk* pk, const zg_fr* const* advice, const zg_fr* const* instances,
                             const size_t* instance_lens, zg_rng_fill_fn rng, void* rng_state, uint8_t* proof_out, size_t proof_cap,
                             size_t* proof_len);

int zg_create_proof(zg_ctx* ctx, zg_pk* pk, const zg_fr* const* advice, const zg_fr* const* instances, const size_t* instance_lens,
                    zg_rng_fill_fn rng, void* rng_state, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
  ZG_ENTER(ctx);
  // The proof runs on the context's own high-priority stream, ordered after whatever the caller has enqueued on
  // ctx->stream (e.g. device-resident advice); it is host-synchronous, so the caller's stream needs no join.
  cudaStream_t caller = ctx->stream;
  ZG_CUDA(cudaEventRecord(ctx->ev_fork, caller));
  ZG_CUDA(cudaStreamWaitEvent(ctx->hp, ctx->ev_fork, 0));
  ctx->stream = ctx->hp;
  int rc = create_proof_impl(ctx, pk, advice, instances, instance_lens, rng, rng_state, proof_out, proof_cap, proof_len);
  ctx->stream = caller;
  // leave nothing in flight, also on the error paths (the workspace is reused by the next proof)
  cudaError_t e1 = cudaStreamSynchronize(ctx->hp), e2 = cudaStreamSynchronize(ctx->aux);
  for (int i = 0; i < zg_ctx::N_SIDE; i++) cudaStreamSynchronize(ctx->side[i]);
  if (rc == ZG_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) return ctx->cuda_fail(e1 != cudaSuccess ? e1 : e2, "create_proof");
  return rc;
}

}  // extern "C"

static int create_proof_impl(zg_ctx* ctx, zg_pk* pk, const zg_fr* const* advice, const zg_fr* const* instances,
                             const size_t* instance_lens, zg_rng_fill_fn rng, void* rng_state, uint8_t* proof_out, size_t proof_cap,
                             size_t* proof_len) {
  if (!pk || !advice || !rng || !proof_out || !proof_len) return ctx->fail(ZG_E_INVALID, "create_proof: null argument");
  // the context's SRS may have been replaced since zg_pk_load: the key is only valid on the bases it was built on
  if (!ctx->srs_loaded || ctx->srs_k != pk->k || !ctx->table[0].pts || !ctx->table[1].pts)
    return ctx->fail(ZG_E_STATE, "create_proof: the context's SRS does not match the proving key (k or a basis missing)");
  if (pk->I && (!instances || !instance_lens)) return ctx->fail(ZG_E_INVALID, "create_proof: instance columns missing");
  for (uint32_t c = 0; c < pk->A; c++)
    if (!advice[c]) return ctx->fail(ZG_E_INVALID, "create_proof: null advice column");
  for (uint32_t c = 0; c < pk->I; c++)
    if (instance_lens[c] && !instances[c]) return ctx->fail(ZG_E_INVALID, "create_proof: null instance column");
  const size_t n = pk->n, N = pk->N;
  const uint32_t A = pk->A, I = pk->I, Lk = pk->n_lookups, S = pk->nsets, bf = pk->bf, usable = pk->usable, m = pk->m;
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  int rc;
  // a proof spread over GPUs (dist.cu): this rank builds the extended forms and the quotient numerator only on the coset
  // blocks c = rank (mod G); the blocks of h are all-gathered before the division by the vanishing polynomial
  const bool dist_rows = ctx->comm && ctx->nranks > 1 && ctx->dist_columns && pk->ext.cosets > 1;
  pk->blk_first = dist_rows ? (uint32_t)ctx->rank : 0;
  pk->blk_step = dist_rows ? (uint32_t)ctx->nranks : 1;
  struct StageEvents {                       // destroyed on every return path
    cudaEvent_t e[9] = {};
    ~StageEvents() { for (auto x : e) if (x) cudaEventDestroy(x); }
  } stage_events;
  cudaEvent_t* ev = stage_events.e;
  for (int i = 0; i < 9; i++) ZG_CUDA(cudaEventCreate(&ev[i]));
  ZG_CUDA(cudaEventRecord(ev[0], st));
  Transcript tr;
  tr.common_scalar(pk->transcript_repr);

  // ---- RNG: every draw of this proof, in upstream order; the small leading draws first -----------------
  const size_t small_draws = (size_t)A * (bf + 1) + A + (size_t)Lk * (2 * (bf + 1) + 2) + (size_t)S * (bf + 1) + (size_t)Lk * (bf + 1);
  rng(rng_state, pk->rnd_words_host, 8 * small_draws);
  ZG_CUDA(cudaMemcpyAsync(pk->rnd_words_dev, pk->rnd_words_host, 64 * small_draws, cudaMemcpyHostToDevice, st));
  fr_from_u512(pk->rnd_words_dev, pk->rnd, small_draws, st, lc);
  bool big_drawn = false;
  auto draw_big = [&]() -> int {           // host draw + upload + conversion of everything after the small draws
    if (big_drawn) return ZG_OK;
    const size_t rest = pk->n_draws - small_draws;
    rng(rng_state, pk->rnd_words_host + 8 * small_draws, 8 * rest);
    ZG_CUDA(cudaMemcpyAsync(pk->rnd_words_dev + 8 * small_draws, pk->rnd_words_host + 8 * small_draws, 64 * rest,
                            cudaMemcpyHostToDevice, st));
    fr_from_u512(pk->rnd_words_dev + 8 * small_draws, pk->rnd + small_draws, rest, st, lc);
    big_drawn = true;
    return ZG_OK;
  };
  size_t draw = 0;  // next unused draw index
  auto blind_rows = [&](Fr* col, size_t first_row, size_t count) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(col + first_row, pk->rnd + draw, sizeof(Fr) * count, cudaMemcpyDeviceToDevice, st);
    draw += count;
    return e;
  };

  // ---- 1. instance columns ------------------------------------------------------------------------------
  ZG_CUDA(cudaMemsetAsync(pk->inst_values, 0, sizeof(Fr) * I * n, st));
  for (uint32_t c = 0; c < I; c++) {
    size_t len = instance_lens ? instance_lens[c] : 0;
    if (len > usable) return ctx->fail(ZG_E_INVALID, "create_proof: instance too large");
    for (size_t i = 0; i < len; i++) {
      Fr v;
      memcpy(v.v, &instances[c][i], 32);
      tr.common_scalar(v);
    }
    if (len) ZG_CUDA(cudaMemcpyAsync(pk->inst_values + c * n, instances[c], sizeof(Fr) * len, cudaMemcpyHostToDevice, st));
  }
  // ---- 2. advice: upload, blind, commit ----------------------------------------------------------------------
  for (uint32_t c = 0; c < A; c++)
    ZG_CUDA(cudaMemcpyAsync(pk->adv_values + c * n, advice[c], sizeof(Fr) * usable, cudaMemcpyDefault, st));  // host or device
  for (uint32_t c = 0; c < A; c++) ZG_CUDA(blind_rows(pk->adv_values + c * n, usable, bf + 1));
  draw += A;  // one Blind(Fr::random) per column, unused by KZG
  {
    // coefficient and extended-coset forms: nobody waits for them before the quotient stage -> low-priority stream
    {
      AuxScope aux(ctx);
      if (aux.err != cudaSuccess) return ctx->cuda_fail(aux.err, "fork to aux stream");
      if (I) {
        rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)pk->inst_values, (zg_fr*)pk->inst_polys, pk->k, I, n);
        if (rc) return rc;
      }
      rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)pk->adv_values, (zg_fr*)pk->adv_polys, pk->k, A, n);
      if (rc) return rc;
      rc = coset_advice_instance(ctx, pk);
      if (rc) return rc;
    }
    std::vector<Affine> aff(A);
    rc = msm_round(ctx, ZG_BASIS_LAGRANGE, pk->adv_values, n, n, A, 0);
    if (rc) return rc;
    // The big random polynomial (n + 1 draws: 8.4 MB at k = 17, 2-3 ms of host time) is drawn while the GPU works: behind
    // the lookup round's kernels when the circuit has lookups (the advice commitment alone is shorter than the draw),
    // here otherwise.  The caller's RNG sees the same call order either way: all small draws, then this one.
    static const bool rng_early = getenv("ZG_RNG_EARLY") != nullptr;      // A/B switch: draw inside the advice round
    if (!Lk || rng_early) {
      int drc = draw_big();
      if (drc) return drc;
    }
    std::vector<G1Jac> jac(A);
    ZG_CUDA(cudaMemcpyAsync(jac.data(), ctx->d_msm_out, sizeof(G1Jac) * A, cudaMemcpyDeviceToHost, st));
    ZG_CUDA(cudaEventRecord(ev[1], st));
    ZG_CUDA(cudaStreamSynchronize(st));
    batch_normalize(jac.data(), A, aff.data());
    for (uint32_t c = 0; c < A; c++)
      if (!tr.write_point(aff[c])) return ctx->fail(ZG_E_SYNTH, "create_proof: advice commitment is the identity");
  }
  const Fr theta = tr.squeeze();

  // ---- 3. lookups: compress, permute, commit --------------------------------------------------------------------
  ExprEnv base_env = make_env(pk, false);
  LookupProgs lp{pk->prog_off, pk->d_in_first, pk->d_in_count, pk->d_tab_first, pk->d_tab_count};
  if (Lk) {
    expr_compress_lookups(base_env, lp, Lk, theta, pk->ci, pk->ct, n, st, lc);
    // Tables with a single expression do not depend on theta: sorted once per proving key (full 256-bit
    // sort) and cached.  The others are sorted per proof on their top 48 bits; the device flags the rare
    // case where that left distinct keys out of order and the round is redone with the full sort.
    const size_t draw_mark = draw;
    std::vector<char> full(Lk, 0), sorted_now(Lk, 0);
    std::vector<Affine> aff(2 * Lk);
    for (int attempt = 0;; attempt++) {
      draw = draw_mark;
      ZG_CUDA(cudaMemsetAsync(pk->d_status, 0, 4 * 2 * Lk, st));
      // the lookups are independent chains of ~40 small kernels each (radix passes, scans, ranking, placement): lookup l
      // runs on side stream l mod 4 with its own workspace, forked from and joined to the proof's stream by events
      ZG_CUDA(cudaEventRecord(ctx->ev_fork, st));
      for (uint32_t l = 0; l < Lk; l++) {
        const int sidx = (int)(l % zg_ctx::N_SIDE);
        cudaStream_t ss = ctx->side[sidx];
        uint8_t* ws = pk->lookup_ws + (size_t)sidx * pk->lookup_ws_stride;
        if (l < (uint32_t)zg_ctx::N_SIDE) ZG_CUDA(cudaStreamWaitEvent(ss, ctx->ev_fork, 0));
        LookupTable tab;
        if (pk->table_cacheable[l]) {
          tab = lookup_table_carve(pk->table_cache + (size_t)l * pk->table_cache_stride, usable);
          if (!pk->table_cached[l]) {
            if (lookup_sort_table(pk->ct + l * n, usable, tab, ws, pk->d_status + 2 * l, true, ss, lc))
              return ctx->cuda_fail(cudaGetLastError(), "lookup_sort_table");
            sorted_now[l] = 1;             // marked cached only once this round has completed without error
          }
        } else {
          tab = lookup_workspace_table(ws, usable);
          if (lookup_sort_table(pk->ct + l * n, usable, tab, ws, pk->d_status + 2 * l, full[l] != 0, ss, lc))
            return ctx->cuda_fail(cudaGetLastError(), "lookup_sort_table");
        }
        if (lookup_permute(pk->ci + l * n, usable, tab, pk->pa + l * n, pk->ps + l * n, ws, pk->d_status + 2 * l + 1, ss, lc))
          return ctx->cuda_fail(cudaGetLastError(), "lookup_permute");
      }
      for (int i = 0; i < zg_ctx::N_SIDE && (uint32_t)i < Lk; i++) {
        ZG_CUDA(cudaEventRecord(ctx->ev_side[i], ctx->side[i]));
        ZG_CUDA(cudaStreamWaitEvent(st, ctx->ev_side[i], 0));
      }
      // blinding rows: per lookup input rows, table rows, then two blinds
      for (uint32_t l = 0; l < Lk; l++) {
        ZG_CUDA(blind_rows(pk->pa + l * n, usable, bf + 1));
        ZG_CUDA(blind_rows(pk->ps + l * n, usable, bf + 1));
        draw += 2;
      }
      {
        AuxScope aux(ctx);
        if (aux.err != cudaSuccess) return ctx->cuda_fail(aux.err, "fork to aux stream");
        rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)pk->pa, (zg_fr*)pk->pa_poly, pk->k, 2 * Lk, n);
        if (rc) return rc;
        rc = coset_lookup_permuted(ctx, pk);
        if (rc) return rc;
      }
      rc = msm_round(ctx, ZG_BASIS_LAGRANGE, pk->pa, n, n, 2 * Lk, 0);
      if (rc) return rc;
      rc = draw_big();                       // host work while the lookup round's kernels run
      if (rc) return rc;
      std::vector<G1Jac> jac(2 * Lk);
      std::vector<uint32_t> status(2 * Lk);
      ZG_CUDA(cudaMemcpyAsync(jac.data(), ctx->d_msm_out, sizeof(G1Jac) * 2 * Lk, cudaMemcpyDeviceToHost, st));
      ZG_CUDA(cudaMemcpyAsync(status.data(), pk->d_status, 4 * 2 * Lk, cudaMemcpyDeviceToHost, st));
      ZG_CUDA(cudaEventRecord(ev[2], st));
      ZG_CUDA(cudaStreamSynchronize(st));
      bool retry = false;
      for (uint32_t l = 0; l < Lk; l++) {
        if (status[2 * l] && !full[l] && !pk->table_cacheable[l]) { full[l] = 1; retry = true; }
      }
      if (retry && attempt == 0) {
        ZG_CUDA(join_aux(ctx));   // the transforms of the failed attempt still read pa / ps
        continue;
      }
      for (uint32_t l = 0; l < Lk; l++)
        if (status[2 * l + 1]) return ctx->fail(ZG_E_SYNTH, "create_proof: lookup input not in table (ConstraintSystemFailure)");
      for (uint32_t l = 0; l < Lk; l++)
        if (sorted_now[l]) pk->table_cached[l] = 1;
      batch_normalize(jac.data(), 2 * Lk, aff.data());
      break;
    }
    // commitments: write order is (input_l, table_l) per lookup
    for (uint32_t l = 0; l < Lk; l++) {
      if (!tr.write_point(aff[l]) || !tr.write_point(aff[Lk + l]))
        return ctx->fail(ZG_E_SYNTH, "create_proof: lookup commitment is the identity");
    }
  } else {
    ZG_CUDA(cudaEventRecord(ev[2], st));
  }
  const Fr beta = tr.squeeze();
  const Fr gamma = tr.squeeze();

  // ---- 4. permutation grand products ------------------------------------------------------------------------------
  // frac[i] = prod_j (v_j + delta^j w^i beta + gamma) / (v_j + beta sigma_j + gamma), via expr-free kernels:
  // numerator and denominator are accumulated with fr_mul_add_scalar-style passes on the host-selected columns.
  auto col_values = [&](uint32_t c) -> const Fr* {
    uint32_t kind = pk->perm[c].first, idx = pk->perm[c].second;
    return kind == 0 ? pk->adv_values + idx * n : kind == 1 ? pk->fixed_values + idx * n : pk->inst_values + idx * n;
  };
  // All S + Lk fraction columns are built first so that ONE batch inversion, ONE multiply and ONE batched
  // scan cover them (these kernels are latency-bound: one Fermat inversion / one block scan per launch).
  // Permutation set s > 0 continues from set s-1's value at row n-(bf+1): its scan starts at 1 and the
  // column is scaled afterwards, which is the same product.
  {
    const uint32_t cnt = S + Lk;
    if (cnt > 16) return ctx->fail(ZG_E_INVALID, "create_proof: too many grand-product columns");
    Fr* num = pk->h;        // cnt columns of n
    Fr* den = pk->h_coeff;
    Fr deltaomega = fr_one();
    for (uint32_t s = 0; s < S; s++) {
      uint32_t c0 = s * pk->chunk, c1 = std::min(c0 + pk->chunk, m);
      std::vector<const Fr*> vals, sigs;
      for (uint32_t c = c0; c < c1; c++) { vals.push_back(col_values(c)); sigs.push_back(pk->sigma_values + c * n); }
      const Fr** dv = pk->d_polyptrs + 2 * (size_t)pk->chunk * s;
      const Fr** ds = dv + vals.size();
      ZG_CUDA(cudaMemcpyAsync(dv, vals.data(), vals.size() * sizeof(Fr*), cudaMemcpyHostToDevice, st));
      ZG_CUDA(cudaMemcpyAsync(ds, sigs.data(), sigs.size() * sizeof(Fr*), cudaMemcpyHostToDevice, st));
      perm_fraction(dv, ds, (uint32_t)vals.size(), pk->omega_pows, beta, gamma, deltaomega, pk->delta, num + s * n, den + s * n, n, st, lc);
      for (uint32_t c = c0; c < c1; c++) deltaomega = fp_mul(deltaomega, pk->delta);
    }
    for (uint32_t l = 0; l < Lk; l++)
      lookup_fraction(pk->ci + l * n, pk->ct + l * n, pk->pa + l * n, pk->ps + l * n, beta, gamma, num + (S + l) * n, den + (S + l) * n, n,
                      st, lc);
    fr_batch_invert(den, (size_t)cnt * n, st, lc, pk->wpoly);       // wpoly (8n) is free until the scan below
    fr_mul_vec(num, den, den, (size_t)cnt * n, st, lc);
    std::vector<uint32_t> nout(cnt);
    for (uint32_t s = 0; s < S; s++) nout[s] = (uint32_t)n;
    for (uint32_t l = 0; l < Lk; l++) nout[S + l] = (uint32_t)(n - bf);
    fr_running_product_batch(den, n, pk->pz, n, nout.data(), cnt, pk->wpoly, st, lc);   // pz | lz are contiguous
    for (uint32_t s = 0; s < S; s++) {
      Fr* z = pk->pz + s * n;
      if (s > 0) fr_scale_by_dev(z, pk->pz + (s - 1) * n + (n - (bf + 1)), n, st, lc);
      ZG_CUDA(blind_rows(z, n - bf, bf));
      draw += 1;
    }
    for (uint32_t l = 0; l < Lk; l++) {
      ZG_CUDA(blind_rows(pk->lz + l * n, n - bf, bf));
      draw += 1;
    }
  }
  // ---- 6. vanishing random polynomial; commit round ------------------------------------------------------------------
  pk->random_poly = pk->rnd + draw;
  draw += n + 1;
  {
    const uint32_t cnt = S + Lk;
    std::vector<G1Jac> jac(cnt + 1);
    if (cnt) {
      AuxScope aux(ctx);
      if (aux.err != cudaSuccess) return ctx->cuda_fail(aux.err, "fork to aux stream");
      rc = zg_lagrange_to_coeff_dev(ctx, (const zg_fr*)pk->pz, (zg_fr*)pk->pz_poly, pk->k, cnt, n);
      if (rc) return rc;
      rc = coset_products(ctx, pk);
      if (rc) return rc;
    }
    // one pipeline for the whole round: the product columns on the Lagrange basis and, as MSM number `cnt`, the random
    // polynomial on the monomial basis (copied behind the product columns so the batch is one strided array)
    ZG_CUDA(cudaMemcpyAsync(pk->pz + (size_t)cnt * n, pk->random_poly, sizeof(Fr) * n, cudaMemcpyDeviceToDevice, st));
    rc = msm_round(ctx, ZG_BASIS_LAGRANGE, pk->pz, n, n, cnt + 1, 1u << cnt);
    if (rc) return rc;
    ZG_CUDA(cudaMemcpyAsync(jac.data(), ctx->d_msm_out, sizeof(G1Jac) * (cnt + 1), cudaMemcpyDeviceToHost, st));
    ZG_CUDA(cudaEventRecord(ev[3], st));
    ZG_CUDA(cudaStreamSynchronize(st));
    std::vector<Affine> aff(cnt + 1);
    batch_normalize(jac.data(), cnt + 1, aff.data());
    for (uint32_t i = 0; i < cnt + 1; i++)
      if (!tr.write_point(aff[i])) return ctx->fail(ZG_E_SYNTH, "create_proof: product commitment is the identity");
  }
  const Fr y = tr.squeeze();

  // ---- 7. quotient numerator on the extended coset ----------------------------------------------------------------------
  ZG_CUDA(join_aux(ctx));   // every coset form is ready
  rc = quotient_numerator(ctx, pk, theta, beta, gamma, y, /*transform=*/false);
  if (rc) return rc;
  if (dist_rows) {
    rc = dist_allgather_blocks(ctx, pk->h, pk->ext.B, pk->ext.cosets);
    if (rc) return rc;
  }
  ZG_CUDA(cudaEventRecord(ev[4], st));
  // ---- 8. vanishing::construct: divide, back to coefficients, commit the pieces ---------------------------------------------
  ext_divide_by_vanishing(pk->ext, pk->h, st, lc);
  rc = ext_to_coeff(ctx, pk->ext, pk->h, pk->adv_cosets /* dead after the numerator */, n * pk->qdeg, pk->h_coeff);
  if (rc) return rc;
  draw += pk->qdeg;  // h piece blinds
  {
    std::vector<Affine> aff(pk->qdeg);
    rc = commit_batch(ctx, ZG_BASIS_MONOMIAL, pk->h_coeff, n, n, pk->qdeg, aff.data(), true);
    if (rc) return rc;
    for (uint32_t i = 0; i < pk->qdeg; i++)
      if (!tr.write_point(aff[i])) return ctx->fail(ZG_E_SYNTH, "create_proof: h commitment is the identity");
  }
  ZG_CUDA(cudaEventRecord(ev[5], st));
  const Fr x = tr.squeeze();
  const Fr xn = fr_pow(x, n);

  // ---- 9. evaluations ------------------------------------------------------------------------------------------------------------
  // h_poly = sum_i xn^i piece_i
  {
    std::vector<const Fr*> pp(pk->qdeg);
    std::vector<Fr> cf(pk->qdeg);
    Fr cur = fr_one();
    for (uint32_t i = 0; i < pk->qdeg; i++) { pp[i] = pk->h_coeff + i * n; cf[i] = cur; cur = fp_mul(cur, xn); }
    ZG_CUDA(cudaMemcpyAsync(pk->d_polyptrs, pp.data(), pp.size() * sizeof(Fr*), cudaMemcpyHostToDevice, st));
    ZG_CUDA(cudaMemcpyAsync(pk->coeff_dev, cf.data(), cf.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));
    fr_linear_combination(pk->d_polyptrs, pk->coeff_dev, pk->qdeg, n, fp_zero<FrParams>(), pk->h_poly, st, lc);
  }
  // queries in create_proof's ProverQuery order
  struct Query { const Fr* poly; int32_t rot; };
  std::vector<Query> Q;
  for (auto& qa : pk->q[0]) Q.push_back({pk->adv_polys + qa.first * n, qa.second});
  const size_t q_perm = Q.size();
  for (uint32_t s = 0; s < S; s++) { Q.push_back({pk->pz_poly + s * n, 0}); Q.push_back({pk->pz_poly + s * n, 1}); }
  for (int s = (int)S - 2; s >= 0; s--) Q.push_back({pk->pz_poly + s * n, -(int32_t)(bf + 1)});
  const size_t q_lk = Q.size();
  for (uint32_t l = 0; l < Lk; l++) {
    Q.push_back({pk->lz_poly + l * n, 0});
    Q.push_back({pk->pa_poly + l * n, 0});
    Q.push_back({pk->ps_poly + l * n, 0});
    Q.push_back({pk->pa_poly + l * n, -1});
    Q.push_back({pk->lz_poly + l * n, 1});
  }
  const size_t q_fixed = Q.size();
  for (auto& qf : pk->q[1]) Q.push_back({pk->fixed_polys + qf.first * n, qf.second});
  const size_t q_sigma = Q.size();
  for (uint32_t c = 0; c < m; c++) Q.push_back({pk->sigma_polys + c * n, 0});
  const size_t q_h = Q.size();
  Q.push_back({pk->h_poly, 0});
  Q.push_back({pk->random_poly, 0});
  // distinct rotations in first-appearance order (construct_intermediate_sets)
  std::vector<int32_t> rots;
  std::vector<uint32_t> pidx(Q.size());
  for (size_t i = 0; i < Q.size(); i++) {
    auto it = std::find(rots.begin(), rots.end(), Q[i].rot);
    if (it == rots.end()) { rots.push_back(Q[i].rot); pidx[i] = (uint32_t)rots.size() - 1; }
    else pidx[i] = (uint32_t)(it - rots.begin());
  }
  if (rots.size() > 16) return ctx->fail(ZG_E_INVALID, "create_proof: more than 16 opening points");
  std::vector<Fr> points(rots.size());
  for (size_t i = 0; i < rots.size(); i++)
    points[i] = fp_mul(x, fr_pow(rots[i] >= 0 ? pk->omega : pk->omega_inv, (uint64_t)(rots[i] >= 0 ? rots[i] : -rots[i])));
  std::vector<const Fr*> qptr(Q.size());
  for (size_t i = 0; i < Q.size(); i++) qptr[i] = Q[i].poly;
  ZG_CUDA(cudaMemcpyAsync(pk->d_polyptrs, qptr.data(), qptr.size() * sizeof(Fr*), cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaMemcpyAsync(pk->d_pidx, pidx.data(), pidx.size() * 4, cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaMemcpyAsync(pk->points_dev, points.data(), points.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));
  fr_eval_many(pk->d_polyptrs, pk->d_pidx, pk->points_dev, (uint32_t)Q.size(), n, pk->evals_dev, pk->scratch + 8 * 4096, st, lc);
  std::vector<Fr> evals(Q.size());
  ZG_CUDA(cudaMemcpyAsync(evals.data(), pk->evals_dev, sizeof(Fr) * Q.size(), cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaEventRecord(ev[6], st));
  ZG_CUDA(cudaStreamSynchronize(st));
  // transcript order: advice, fixed, random, sigma, permutation sets, lookups
  for (size_t i = 0; i < q_perm; i++) tr.write_scalar(evals[i]);
  for (size_t i = q_fixed; i < q_sigma; i++) tr.write_scalar(evals[i]);
  tr.write_scalar(evals[q_h + 1]);
  for (size_t i = q_sigma; i < q_h; i++) tr.write_scalar(evals[i]);
  for (uint32_t s = 0; s < S; s++) {
    tr.write_scalar(evals[q_perm + 2 * s]);
    tr.write_scalar(evals[q_perm + 2 * s + 1]);
    if (s + 1 < S) tr.write_scalar(evals[q_perm + 2 * S + (S - 2 - s)]);
  }
  for (uint32_t l = 0; l < Lk; l++) {
    const size_t b = q_lk + 5 * l;
    tr.write_scalar(evals[b + 0]);   // product
    tr.write_scalar(evals[b + 4]);   // product at omega x
    tr.write_scalar(evals[b + 1]);   // permuted input
    tr.write_scalar(evals[b + 3]);   // permuted input at omega^-1 x
    tr.write_scalar(evals[b + 2]);   // permuted table
  }
  // ---- 10. GWC multi-open -------------------------------------------------------------------------------------------------------------
  const Fr v = tr.squeeze();
  const uint32_t nsetsQ = (uint32_t)rots.size();
  if (nsetsQ > 8) return ctx->fail(ZG_E_INVALID, "create_proof: more than 8 opening sets");
  ZG_CUDA(cudaMemsetAsync(pk->wpoly, 0, sizeof(Fr) * nsetsQ * n, st));
  std::vector<const Fr*> gp;
  std::vector<Fr> gc;
  std::vector<Fr> eaccs(nsetsQ);
  std::vector<size_t> goff(nsetsQ + 1, 0);
  for (uint32_t g = 0; g < nsetsQ; g++) {
    Fr pw = fr_one(), eacc = fp_zero<FrParams>();
    for (size_t i = 0; i < Q.size(); i++) {
      if (pidx[i] != g) continue;
      gp.push_back(Q[i].poly);
      gc.push_back(pw);
      eacc = fp_add(eacc, fp_mul(evals[i], pw));
      pw = fp_mul(pw, v);
    }
    eaccs[g] = eacc;
    goff[g + 1] = gp.size();
  }
  ZG_CUDA(cudaMemcpyAsync(pk->d_polyptrs, gp.data(), gp.size() * sizeof(Fr*), cudaMemcpyHostToDevice, st));
  ZG_CUDA(cudaMemcpyAsync(pk->coeff_dev, gc.data(), gc.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));
  {
    GwcBatch gb;
    gb.nsets = nsetsQ;
    for (uint32_t g = 0; g <= nsetsQ; g++) gb.off[g] = (uint32_t)goff[g];
    for (uint32_t g = 0; g < nsetsQ; g++) { gb.sub0[g] = eaccs[g]; gb.z[g] = points[g]; }
    fr_gwc_witness_batch(pk->d_polyptrs, pk->coeff_dev, gb, n, pk->fold, pk->wpoly, pk->scratch, st, lc);
  }
  {
    std::vector<Affine> aff(nsetsQ);
    rc = commit_batch(ctx, ZG_BASIS_MONOMIAL, pk->wpoly, n, n, nsetsQ, aff.data(), true);
    if (rc) return rc;
    for (uint32_t g = 0; g < nsetsQ; g++)
      if (!tr.write_point(aff[g])) return ctx->fail(ZG_E_SYNTH, "create_proof: opening witness is the identity");
  }
  ZG_CUDA(cudaEventRecord(ev[7], st));
  ZG_CUDA(cudaEventSynchronize(ev[7]));
  for (int i = 0; i < 7; i++) cudaEventElapsedTime(&pk->stage_ms[i], ev[i], ev[i + 1]);
  cudaEventElapsedTime(&pk->stage_ms[7], ev[0], ev[7]);
  if (draw != pk->n_draws) return ctx->fail(ZG_E_STATE, "create_proof: RNG draw accounting mismatch");
  if (tr.out.size() > proof_cap) return ctx->fail(ZG_E_INVALID, "create_proof: proof buffer too small");
  memcpy(proof_out, tr.out.data(), tr.out.size());
  *proof_len = tr.out.size();
  return ZG_OK;
}

extern "C" {

// plonk::evaluation::Evaluator::evaluate_h for one circuit: polynomials in coefficient form (host), result on the
// extended coset (host).  divide != 0 also applies EvaluationDomain::divide_by_vanishing_poly.
int zg_evaluate_h(zg_ctx* ctx, zg_pk* pk, const zg_fr* const* advice_polys, const zg_fr* const* instance_polys,
                  const zg_fr* const* lookup_input_polys, const zg_fr* const* lookup_table_polys,
                  const zg_fr* const* lookup_product_polys, const zg_fr* const* perm_product_polys, const zg_fr challenges[4],
                  int divide, zg_fr* h_out) {
  ZG_ENTER(ctx);
  if (!pk || !challenges || !h_out) return ctx->fail(ZG_E_INVALID, "evaluate_h: null argument");
  const size_t n = pk->n, N = pk->N;
  const uint32_t A = pk->A, I = pk->I, Lk = pk->n_lookups, S = pk->nsets;
  if ((A && !advice_polys) || (I && !instance_polys) || (S && !perm_product_polys) ||
      (Lk && (!lookup_input_polys || !lookup_table_polys || !lookup_product_polys)))
    return ctx->fail(ZG_E_INVALID, "evaluate_h: missing polynomial list");
  cudaStream_t st = ctx->stream;
  LaunchCounter lc{&ctx->launches};
  auto up = [&](Fr* dst, const zg_fr* const* src, uint32_t count) -> cudaError_t {
    for (uint32_t c = 0; c < count; c++) {
      if (!src[c]) return cudaErrorInvalidValue;
      cudaError_t e = cudaMemcpyAsync(dst + c * n, src[c], sizeof(Fr) * n, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  ZG_CUDA(up(pk->adv_polys, advice_polys, A));
  ZG_CUDA(up(pk->inst_polys, instance_polys, I));
  ZG_CUDA(up(pk->pz_poly, perm_product_polys, S));
  ZG_CUDA(up(pk->pa_poly, lookup_input_polys, Lk));
  ZG_CUDA(up(pk->ps_poly, lookup_table_polys, Lk));
  ZG_CUDA(up(pk->lz_poly, lookup_product_polys, Lk));
  Fr ch[4];
  memcpy(ch, challenges, sizeof(ch));   // theta, beta, gamma, y
  pk->blk_first = 0;
  pk->blk_step = 1;
  int rc = quotient_numerator(ctx, pk, ch[0], ch[1], ch[2], ch[3], /*transform=*/true);
  if (rc) return rc;
  // The numerator lives on the internal domain (extdomain.cuh).  The ABI speaks halo2's coset zeta * <omega_ext>:
  // quotient -> coefficients -> values on that coset, times (X^n - 1) again when the caller wants the numerator.
  ext_divide_by_vanishing(pk->ext, pk->h, st, lc);
  const size_t keep = std::min(pk->N, n * pk->qdeg);
  rc = ext_to_coeff(ctx, pk->ext, pk->h, pk->adv_cosets, keep, pk->h_coeff);
  if (rc) return rc;
  const size_t NH = (size_t)1 << pk->ext_k;
  rc = ws_reserve(ctx, ctx->ws_stage, sizeof(Fr) * NH);
  if (rc) return rc;
  Fr* hh = (Fr*)ctx->ws_stage.p;
  rc = ntt_halo_coset_dev(ctx, pk->h_coeff, (uint32_t)keep, pk->ext_k, hh);
  if (rc) return rc;
  if (!divide) fr_mul_periodic(hh, pk->t_inv, 1u << (pk->ext_k - pk->k), NH, st, lc);
  ZG_CUDA(cudaMemcpyAsync(h_out, hh, sizeof(Fr) * NH, cudaMemcpyDeviceToHost, st));
  ZG_CUDA(cudaStreamSynchronize(st));
  return ZG_OK;
}

}  // extern "C"
