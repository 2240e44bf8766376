/* zg_b200 -- C ABI of the B200-native halo2/KZG proving backend for BN254.
 *
 * This is the drop-in boundary for the proving hot path that `zero_g` reaches through
 * halo2_proofs (tag v2023_04_20) from /root/reference/src/wnn.rs:226-228 (keygen_vk/keygen_pk)
 * and :242-259 (create_proof).  halo2_proofs has no plugin interface; the seams are its free
 * functions / inherent methods, and each entry point below names the one it replaces.  A
 * maintainer binds these from a patched halo2_proofs (Cargo `[patch]`, the mechanism the
 * reference already uses at Cargo.toml:14-18); see INTEGRATION.md for the Rust `extern "C"` stub.
 *
 * Conventions
 *  - every function returns 0 (ZG_OK) or a negative ZG_E_* code and never throws or aborts
 *    across the boundary; zg_last_error(ctx) returns a message for the last failure.
 *  - zg_fr / zg_fq: 4 x u64 little-endian limbs in Montgomery form (R = 2^256): the in-memory
 *    layout of halo2curves 0.3.3 `bn256::Fr` / `Fq`, so Rust passes `&[Fr]` unconverted.
 *  - zg_g1_affine = {x, y} (identity = (0,0)) = halo2curves `G1Affine`; zg_g1 = Jacobian
 *    {x, y, z} (identity z = 0) = halo2curves `G1`.  A zg_g1 result is a valid representative of
 *    the mathematically unique group element; only its affine normalisation is canonical.
 *  - plain names take HOST pointers and include the host<->device copies; `_dev` names take
 *    DEVICE pointers on the context's device and enqueue on the context's stream without
 *    synchronising (the caller synchronises with zg_sync or its own stream primitives).
 *  - a zg_ctx is bound to one CUDA device and one stream, is NOT thread-safe, and owns all the
 *    device memory it allocates.  Distinct contexts are independent (one per GPU / per process).
 *  - there is no CPU fallback: without a CUDA device zg_ctx_create fails with ZG_E_CUDA.
 */
#ifndef ZG_B200_H
#define ZG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zg_ctx zg_ctx;

typedef struct { uint64_t l[4]; } zg_fr;
typedef struct { uint64_t l[4]; } zg_fq;
typedef struct { zg_fq x, y; } zg_g1_affine;
typedef struct { zg_fq x, y, z; } zg_g1;
/* G2Affine on the twist y^2 = x^3 + 3/(9+u) over Fq2 = Fq[u]/(u^2+1): x = x_c0 + x_c1 u, y likewise (the in-memory
 * layout of halo2curves' bn256::G2Affine); identity = all zero */
typedef struct { zg_fq x_c0, x_c1, y_c0, y_c1; } zg_g2_affine;

enum {
  ZG_OK = 0,
  ZG_E_INVALID = -1,  /* bad argument */
  ZG_E_CUDA = -2,     /* CUDA runtime error (message in zg_last_error) */
  ZG_E_STATE = -3,    /* required object (SRS, pk) not loaded */
  ZG_E_NOMEM = -4,
  ZG_E_SYNTH = -5,    /* prover-level failure, e.g. lookup input not in table (plonk::Error) */
  ZG_E_VERIFY = -6,   /* zg_verify_proof: the proof was rejected (plonk::Error::ConstraintSystemFailure / Opening) */
};

enum { ZG_BASIS_MONOMIAL = 0, ZG_BASIS_LAGRANGE = 1 };

/* ---- context ------------------------------------------------------------------------- */
/* `stream` is a cudaStream_t to enqueue on (e.g. the caller's current stream) or NULL to let the
 * context create its own non-blocking stream. */
int zg_ctx_create(int device, void* stream, zg_ctx** out);
void zg_ctx_destroy(zg_ctx* ctx);
const char* zg_last_error(const zg_ctx* ctx);
int zg_sync(zg_ctx* ctx);
/* number of kernels this context has launched so far (for bench.py's `gpu_launches`) */
uint64_t zg_launch_count(const zg_ctx* ctx);
const char* zg_version(void);

/* raw device memory owned by the caller, on the context's device */
int zg_dev_alloc(zg_ctx* ctx, size_t bytes, void** out);
int zg_dev_free(zg_ctx* ctx, void* p);
int zg_h2d(zg_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int zg_d2h(zg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- SRS: halo2_proofs ParamsKZG<Bn256>::{g, g_lagrange}  (src/main.rs:232, src/io.rs:139-146)
 * Uploads both bases (n = 2^k points each; either may be NULL) and builds the fixed-base window
 * tables used by every MSM on that basis.  One SRS per context. */
int zg_srs_load(zg_ctx* ctx, uint32_t k, const zg_g1_affine* g, const zg_g1_affine* g_lagrange);
/* `ctx` uses the parameters `from` has loaded (both on the same device): no upload, no second set of window tables; the
 * memory is freed with the last context that uses it.  For several contexts that prove side by side on one GPU. */
int zg_srs_share(zg_ctx* ctx, zg_ctx* from);

/* ---- MSM: arithmetic::best_multiexp via ParamsKZG::commit (basis 0) / commit_lagrange (1) --- */
/* out = sum_i scalars[i] * basis[i], n <= 2^k */
int zg_msm(zg_ctx* ctx, int basis, const zg_fr* scalars, size_t n, zg_g1* out);
/* `count` MSMs of n scalars each (one Fiat-Shamir round's commitments) in one pipeline */
int zg_msm_batch(zg_ctx* ctx, int basis, const zg_fr* const* scalars, size_t n, size_t count,
                 zg_g1* out);
/* device-resident scalars: polynomial j starts at scalars_dev + j*stride (elements);
 * out_dev receives `count` zg_g1 */
int zg_msm_dev(zg_ctx* ctx, int basis, const zg_fr* scalars_dev, size_t stride, size_t n,
               size_t count, zg_g1* out_dev);

/* ---- NTT: arithmetic::best_fft and the EvaluationDomain transforms ---------------------- */
/* in place, natural order in and out, a[i] <- sum_j a[j] * omega^(ij); n = 2^log_n */
int zg_ntt(zg_ctx* ctx, zg_fr* a, uint32_t log_n, const zg_fr* omega);
/* `batch` transforms, polynomial j at in_dev + j*stride -> out_dev + j*stride (in == out allowed) */
int zg_ntt_dev(zg_ctx* ctx, const zg_fr* in_dev, zg_fr* out_dev, uint32_t log_n, const zg_fr* omega,
               size_t batch, size_t stride);
/* EvaluationDomain::lagrange_to_coeff: inverse NTT over the 2^k domain (scaled by 1/n) */
int zg_lagrange_to_coeff(zg_ctx* ctx, zg_fr* a, uint32_t k);
int zg_lagrange_to_coeff_dev(zg_ctx* ctx, const zg_fr* in_dev, zg_fr* out_dev, uint32_t k,
                             size_t batch, size_t stride);
/* EvaluationDomain::coeff_to_extended: coeff (2^k) -> evaluations over the coset zeta*<omega_ext>
 * of size 2^ext_k (zeta = Fr::ZETA, as upstream's distribute_powers_zeta) */
int zg_coeff_to_extended(zg_ctx* ctx, const zg_fr* coeff, uint32_t k, uint32_t ext_k, zg_fr* out);
int zg_coeff_to_extended_dev(zg_ctx* ctx, const zg_fr* coeff_dev, size_t in_stride, uint32_t k,
                             uint32_t ext_k, zg_fr* out_dev, size_t out_stride, size_t batch);
/* EvaluationDomain::extended_to_coeff: inverse of the above; writes the first `keep` coefficients */
int zg_extended_to_coeff(zg_ctx* ctx, const zg_fr* ext, uint32_t k, uint32_t ext_k, size_t keep,
                         zg_fr* out);
int zg_extended_to_coeff_dev(zg_ctx* ctx, const zg_fr* ext_dev, uint32_t k, uint32_t ext_k,
                             size_t keep, zg_fr* out_dev);

/* ---- keygen + create_proof ----------------------------------------------------------------
 * plonk::{keygen_vk, keygen_pk} (src/wnn.rs:226-228) and plonk::create_proof::<KZGCommitmentScheme<Bn256>,
 * ProverGWC, _, _, EvmTranscript, _> (src/wnn.rs:242-259) for ONE circuit.  Witness synthesis stays on
 * the host: the caller passes the synthesized columns.  The transcript (keccak256 EvmTranscript) and the
 * RNG draw order are those of halo2_proofs v2023_04_20. */
typedef struct zg_pk zg_pk;

/* RngCore::next_u64 in bulk: fill out[0..n) with the next n u64 draws of the caller's RNG */
typedef void (*zg_rng_fill_fn)(void* state, uint64_t* out, size_t n);
/* rand_xorshift::XorShiftRng restated (seedable RNG for tests / benches; the reference uses OsRng) */
typedef struct { uint32_t x, y, z, w; } zg_xorshift;
void zg_xorshift_seed(zg_xorshift* rng, const uint8_t seed[16]);
void zg_xorshift_fill(void* state /* zg_xorshift* */, uint64_t* out, size_t n);
/* The production RNG: ChaCha20 keystream (RFC 8439 block function, 64-bit block counter) keyed from the operating
 * system's entropy source -- what rand::rngs::OsRng gives the reference at /root/reference/src/wnn.rs:256 (every
 * blinding row, the random polynomial and the h blinds of a proof come from it).  zg_chacha20_seed_os draws a fresh
 * 256-bit key with getrandom(2); zg_chacha20_seed takes a caller-supplied key (known-answer tests). */
typedef struct { uint32_t key[8]; uint64_t counter; uint32_t nonce[2]; uint32_t have; uint8_t buf[64]; } zg_chacha20;
int  zg_chacha20_seed_os(zg_chacha20* rng);                       /* ZG_OK, or ZG_E_STATE when the OS gives no entropy */
void zg_chacha20_seed(zg_chacha20* rng, const uint8_t key[32]);
void zg_chacha20_fill(void* state /* zg_chacha20* */, uint64_t* out, size_t n);

typedef struct {
  uint32_t k;
  /* constraint system after selector compression, serialised by the host front-end
   * (0g-halo2_b200/zg_b200/plonk/serialize.py documents the word stream) */
  const uint32_t* cs_words; size_t cs_nwords;
  const zg_fr* constants; size_t n_constants;    /* constant pool of the expression programs */
  const zg_fr* const* fixed;                     /* num_fixed columns of 2^k Lagrange values */
  const uint32_t* perm_mapping;                  /* [m][2^k][2]: (column, row) -> (column', row') of sigma; or NULL ... */
  zg_fr transcript_repr;                         /* VerifyingKey::transcript_repr, hashed first */
  const zg_fr* const* sigma_values;              /* ... with the m permutation polynomials' Lagrange values given instead
                                                  * (permutation::ProvingKey::permutations of a serialized key,
                                                  * /root/reference/src/io.rs:166-170 read_pk); NULL when perm_mapping is set */
} zg_pk_desc;

/* keygen: commits fixed and sigma columns, builds coefficient + extended-coset forms, l_0/l_last/
 * l_active; everything stays resident on the context's device.  Needs the SRS (zg_srs_load, same k). */
int zg_pk_load(zg_ctx* ctx, const zg_pk_desc* desc, zg_pk** out);
void zg_pk_free(zg_ctx* ctx, zg_pk* pk);
/* VerifyingKey::transcript_repr is a hash over the vk, which contains the commitments zg_pk_load has just computed: a
 * caller that derives it from zg_pk_commitments sets it here instead of loading the key a second time */
int zg_pk_set_transcript_repr(zg_ctx* ctx, zg_pk* pk, const zg_fr* transcript_repr);
/* VerifyingKey: fixed_commitments (num_fixed) and permutation commitments (m), affine */
int zg_pk_commitments(zg_ctx* ctx, const zg_pk* pk, zg_g1_affine* fixed_out, zg_g1_affine* sigma_out);
/* Columns of a resident key, copied to the host (2^k elements): what `ProvingKey::write` serialises
 * (/root/reference/src/io.rs:159-163 write_keys).  The extended-domain forms of the file format are halo2's
 * (zeta coset of size 2^extended_k) and are rebuilt by the caller with zg_coeff_to_extended. */
enum { ZG_PK_FIXED_VALUES = 0, ZG_PK_FIXED_POLYS = 1, ZG_PK_SIGMA_VALUES = 2, ZG_PK_SIGMA_POLYS = 3 };
int zg_pk_read_column(zg_ctx* ctx, const zg_pk* pk, int what, uint32_t index, zg_fr* out);
/* A second key over the SAME resident columns (fixed / sigma forms, domain tables, programs: 672 MiB at k = 17 that are
 * read-only after zg_pk_load): own per-proof workspace, usable from another context of the same device whose SRS has the
 * same k.  Both keys are freed with zg_pk_free; the shared half goes with the last one. */
int zg_pk_clone(zg_ctx* ctx, const zg_pk* src, zg_pk** out);

/* advice: num_advice columns of 2^k values, host pointers (or device pointers on the context's device: the
 * copy is direction-agnostic); rows >= 2^k - blinding_factors - 1 are replaced by blinding scalars; instances:
 * num_instance host columns with their lengths.  Writes the proof bytes.
 * Host-synchronous: the proof is ordered after the work already enqueued on the context's stream, runs on two
 * context-owned streams (critical path at the highest priority, coefficient / extended-coset transforms at the lowest)
 * and has completed when the call returns.  Contexts are independent: one host thread per context may prove
 * concurrently on the same GPU (each with its own SRS tables and zg_pk). */
int zg_create_proof(zg_ctx* ctx, zg_pk* pk, const zg_fr* const* advice, const zg_fr* const* instances,
                    const size_t* instance_lens, zg_rng_fill_fn rng, void* rng_state, uint8_t* proof_out,
                    size_t proof_cap, size_t* proof_len);
/* device time (ms) of the stages of the last zg_create_proof on this pk, for profiling:
 * out[0..8) = advice, lookups, products, quotient, h-commit, evals, gwc, total */
int zg_pk_last_stage_ms(const zg_pk* pk, float out[8]);

/* ---- single prover stages (host pointers) --------------------------------------------------------
 * zg_create_proof keeps a whole proof on the device.  A halo2_proofs fork that replaces ONE upstream function
 * at a time binds these; each wraps exactly the kernels the proof path runs for that step.  Upstream names are
 * halo2_proofs v2023_04_20 (un-vendored, /root/reference/Cargo.toml:21-25), all reached from create_proof
 * (/root/reference/src/wnn.rs:242-259). */
/* plonk::lookup::prover::permute_expression_pair without the blinding rows: a / s are the compressed input /
 * table expressions over the first `usable` rows.  ZG_E_SYNTH when an input is not in the table
 * (upstream Error::ConstraintSystemFailure). */
int zg_lookup_permute(zg_ctx* ctx, const zg_fr* a, const zg_fr* s, size_t usable, zg_fr* a_perm, zg_fr* s_perm);
/* grand product of plonk::permutation::prover::commit / plonk::lookup::prover::commit_product:
 * z[0] = 1, z[i] = z[i-1] * num[i-1] / den[i-1] for i < len (one batch inversion of den) */
int zg_grand_product(zg_ctx* ctx, const zg_fr* num, const zg_fr* den, size_t len, zg_fr* z);
/* ff::BatchInvert::batch_invert in place; zeros stay zero */
int zg_batch_invert(zg_ctx* ctx, zg_fr* a, size_t n);
/* arithmetic::eval_polynomial: out[j] = polys[j](x), every polynomial has n coefficients */
int zg_eval_poly_batch(zg_ctx* ctx, const zg_fr* const* polys, size_t n, size_t count, const zg_fr* x, zg_fr* out);
/* arithmetic::kate_division: q(X) = (a(X) - a(z)) / (X - z); a has n coefficients, q receives n - 1 */
int zg_kate_division(zg_ctx* ctx, const zg_fr* a, size_t n, const zg_fr* z, zg_fr* q);
/* plonk::evaluation::Evaluator::evaluate_h for one circuit of `pk`: polynomials in coefficient form (2^k each):
 * advice (num_advice), instance (num_instance), per lookup the permuted input / permuted table / product
 * polynomials, per permutation set the product polynomial; challenges = {theta, beta, gamma, y}.
 * h_out receives the 2^ext_k values of the y-folded numerator on the extended coset; with divide != 0 it is also
 * multiplied by 1/(X^n - 1) (EvaluationDomain::divide_by_vanishing_poly).  Uses pk's per-proof workspace. */
int zg_evaluate_h(zg_ctx* ctx, zg_pk* pk, const zg_fr* const* advice_polys, const zg_fr* const* instance_polys,
                  const zg_fr* const* lookup_input_polys, const zg_fr* const* lookup_table_polys,
                  const zg_fr* const* lookup_product_polys, const zg_fr* const* perm_product_polys,
                  const zg_fr challenges[4], int divide, zg_fr* h_out);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink / NVSwitch) ---------------------------------------------
 * SURVEY.md 8(e).  A context joins a communicator with zg_comm_init (rank 0 creates the id with zg_comm_unique_id and
 * the host plumbing -- torch.distributed, MPI, a file -- hands its 128 bytes to the other ranks).  NCCL is loaded with
 * dlopen at the first use; single-GPU users never touch it.
 *  - zg_msm_sharded / _dev: ONE large MSM split by point range (BASELINE configs[3], [4]): the context's SRS holds this
 *    rank's slice of the bases (zg_srs_load with the slice), `scalars` the matching slice; the G partial sums (96 B each)
 *    are all-gathered and added on the device; every rank receives the total.
 *  - zg_ctx_set_distribution(ZG_DIST_COLUMNS): zg_create_proof spreads ONE proof over the ranks: the commitments of every
 *    Fiat-Shamir round by column (column j -> rank j mod G, the 96-byte points all-gathered), the extended forms and the
 *    quotient numerator by coset block of the internal extended domain (block c -> rank c mod G, the blocks of h
 *    all-gathered).  Every rank must call zg_create_proof with the same key, witness, instances and RNG stream (SPMD);
 *    every rank returns the same proof bytes. */
enum { ZG_DIST_NONE = 0, ZG_DIST_COLUMNS = 1 };
int zg_comm_unique_id(uint8_t out[128]);
int zg_comm_init(zg_ctx* ctx, int nranks, int rank, const uint8_t unique_id[128]);
int zg_comm_destroy(zg_ctx* ctx);
int zg_ctx_set_distribution(zg_ctx* ctx, int mode);
int zg_msm_sharded_dev(zg_ctx* ctx, int basis, const zg_fr* scalars_dev, size_t stride, size_t n_local, size_t count,
                       zg_g1* out_dev);
int zg_msm_sharded(zg_ctx* ctx, int basis, const zg_fr* scalars, size_t n_local, zg_g1* out);

/* ---- verifier (host code, no GPU) ------------------------------------------------------------------------------
 * plonk::verify_proof::<KZGCommitmentScheme<Bn256>, VerifierGWC, _, EvmTranscript, SingleStrategy> as called by
 * `Wnn::verify_proof` (/root/reference/src/wnn.rs:265-280; `bench_verification`, benches/bench.rs:38-45).  Verification
 * is host work in the reference too; it needs no context and no device.  A zg_vk holds the VerifyingKey data: the
 * constraint system (same blob as zg_pk_desc), the fixed / permutation commitments and transcript_repr.
 * zg_verify_proof returns ZG_OK when the proof is accepted, ZG_E_VERIFY when it is rejected (malformed encoding, a
 * point off the curve, a failed identity or pairing check), ZG_E_INVALID for bad arguments.  `g1_generator` is
 * ParamsKZG::get_g()[0]; `g2`, `s_g2` are ParamsKZG::g2() and ::s_g2().  Instances as for zg_create_proof.
 * zg_pairing_check: is prod_i e(p_i, q_i) the identity of GT (the check behind DualMSM::check)? */
typedef struct zg_vk zg_vk;
int zg_vk_create(uint32_t k, const uint32_t* cs_words, size_t cs_nwords, const zg_fr* constants, size_t n_constants,
                 const zg_g1_affine* fixed_commitments, const zg_g1_affine* permutation_commitments,
                 const zg_fr* transcript_repr, zg_vk** out);
void zg_vk_free(zg_vk* vk);
const char* zg_vk_last_error(const zg_vk* vk);
int zg_verify_proof(zg_vk* vk, const zg_g1_affine* g1_generator, const zg_g2_affine* g2, const zg_g2_affine* s_g2,
                    const zg_fr* const* instances, const size_t* instance_lens, const uint8_t* proof, size_t proof_len);
int zg_pairing_check(const zg_g1_affine* p, const zg_g2_affine* q, size_t n, int* is_one);

/* ---- native witness synthesis of the WNN circuit (host code, no GPU) ----------------------------------------
 * The assignment half of `WnnCircuit::synthesize` -> `WnnChip::predict` (/root/reference/src/gadgets/wnn.rs:180-237,
 * 372-393) and every sub-chip, under halo2's SimpleFloorPlanner: produces the six advice columns zg_create_proof takes.
 * In the reference this is Rust on the host; here it is the native counterpart of zg_b200/plonk/gadgets.py, checked
 * cell for cell against it (tests/test_wnn_synth.py).  A zg_wnn is immutable after creation and may be shared by
 * threads; each zg_wnn_synthesize call writes only its own output buffers (zg_wnn_last_error is per calling thread). */
typedef struct zg_wnn zg_wnn;
typedef struct {
  uint64_t p;                        /* hash modulus (Wnn::p) */
  uint32_t n_hashes, bits_per_hash;  /* WnnCircuitParams::{n_hashes, bits_per_hash}; l = n_hashes * bits_per_hash */
  uint32_t bits_per_filter;          /* inputs per bloom filter */
  uint32_t n_classes, n_filters;     /* bloom_filters.shape[0..2] */
  uint32_t width, height, bits_per_input;
  const uint16_t* thresholds;        /* [width][height][bits_per_input], 0..256 (src/io.rs:59-73) */
  const uint64_t* input_permutation; /* width*height*bits_per_input entries */
  const uint8_t* bloom_bits;         /* [n_classes][n_filters][2^bits_per_hash], 0 / 1 */
} zg_wnn_desc;
int zg_wnn_create(const zg_wnn_desc* desc, zg_wnn** out);
void zg_wnn_free(zg_wnn* w);
const char* zg_wnn_last_error(const zg_wnn* w);
/* image: width*height bytes; advice: 6 columns of 2^k elements (zeroed, then filled below `usable_rows` = 2^k -
 * blinding_factors - 1); outputs: n_classes class scores (the public instance).  ZG_E_SYNTH when the circuit does
 * not fit in `usable_rows` rows (plonk::Error::NotEnoughRowsAvailable). */
int zg_wnn_synthesize(const zg_wnn* w, const uint8_t* image, uint32_t k, uint32_t usable_rows, zg_fr* const* advice,
                      uint64_t* outputs);

/* ---- micro-benchmarks used for the integer-pipe roofline (bench.py) ---------------------- */
/* runs `iters` dependent-free IMAD-class instructions per thread on every SM and returns the
 * achieved rate in 1e9 thread-instructions per second; kind 0 = IMAD (32-bit), 1 = IMAD.WIDE,
 * 2 = Fr Montgomery multiplications, portable body (result in 1e9 mulmod/s), 3 = same, row-wise PTX
 * carry-chain body, 4 = same, even/odd carry-chain body (the one the kernels use); 5 = DFMA (FP64 pipe), 6 = DFMA and
 * IMAD.WIDE interleaved 1:1 (rate counts both), 7 = IMAD.WIDE with 16 independent accumulators and nothing else,
 * 8 = Fr squarings, dedicated squaring body, 9 = Fr two-product multiply-adds with one reduction (rate counts BOTH
 * products of a call; csrc/field_gen.cuh) */
int zg_bench_int_pipe(zg_ctx* ctx, int kind, uint32_t iters, double* giga_per_s);

/* live timing of the dominant kernel (msm_accumulate_kernel, the level-0 bucket accumulation of every MSM): while
 * enabled, each launch is bracketed by CUDA events on the launching stream and its entry count (= mixed point
 * additions) is copied back.  zg_probe_read returns the totals since the last enable / read and resets them. */
int zg_probe_enable(zg_ctx* ctx, int on);
int zg_probe_read(zg_ctx* ctx, double* kernel_ms, uint64_t* launches, uint64_t* point_additions);

/* ---- diagnostics (parity tests) --------------------------------------------------------- */
/* element-wise device field op on host arrays of n elements; field 0 = Fr, 1 = Fq;
 * op 0 mul, 1 mul (portable body), 2 mul (row-wise PTX body), 3 add, 4 sub, 5 inverse(a), 6 from_mont(a),
 * 7 to_mont(a), 8 mul (even/odd carry-chain body), 9 a^2 (dedicated squaring body), 10 a*b + (a+b)*a and
 * 11 a^2 - b^2 (two-product body with one Montgomery reduction; csrc/field_gen.cuh) */
int zg_debug_field_op(zg_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n);
/* keccak256 as used by the EvmTranscript inside zg_create_proof (host code; needs no context and no GPU) */
void zg_debug_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif /* ZG_B200_H */
