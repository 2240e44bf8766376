/* zg_oracle.c -- CPU restatement of the halo2 proving arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libzg_b200.so) never links or calls it.
 *
 * The algorithms restated here live in third-party crates that /root/reference pins but does
 * not vendor (Cargo.toml:14-28): halo2_proofs tag v2023_04_20 (`arithmetic::{best_multiexp,
 * multiexp_serial, best_fft, recursive_butterfly_arithmetic, eval_polynomial, kate_division}`,
 * `poly::EvaluationDomain`) and halo2curves tag 0.3.3 (`bn256::{Fr,Fq,G1}` 4x64-bit Montgomery).
 * They are reached from /root/reference/src/wnn.rs:226-228 and :242-259.  The reference cannot
 * be compiled here (no Rust toolchain), so this is a "port" baseline, and PARITY IS UNPINNED
 * against the real crate: what pins this file are mathematical identities (tests/test_oracle.py:
 * naive DFT, double-and-add MSM, Python big-int field arithmetic).
 *
 * Threading mirrors upstream's rayon use with OpenMP: best_multiexp splits the points into one
 * contiguous chunk per thread and runs the serial Pippenger (window c = ceil(ln n)) on each;
 * best_fft bit-reverses, precomputes n/2 twiddles serially, then recurses on halves in parallel.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;           /* Montgomery form, R = 2^256 */
typedef struct { fe x, y; } g1a;                /* affine, identity = (0,0) */
typedef struct { fe x, y, z; } g1j;             /* Jacobian, identity z = 0 */

typedef struct { uint64_t p[4]; uint64_t inv; fe r, r2; } field_t;

static const field_t FR = {
  {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
  0xc2e1f593efffffffull,
  {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
  {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}}};
static const field_t FQ = {
  {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
  0x87d20782e4866389ull,
  {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
  {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}}};

/* ---- field (halo2curves src/bn256/{fr,fq}.rs via field_arithmetic! macro: mul = schoolbook
 * 4x4 then montgomery_reduce; add/sub with conditional correction) ---- */
static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
  return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline void fe_cond_sub(fe* r, const uint64_t t[4], const field_t* F) {
  uint64_t u[4];
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)t[i] - F->p[i] - (uint64_t)bw;
    u[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  const uint64_t* s = bw ? t : u;
  for (int i = 0; i < 4; i++) r->l[i] = s[i];
}
static inline void fe_add(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[4];
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  fe_cond_sub(r, t, F);
}
static inline void fe_sub(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[4];
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)bw;
    t[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)t[i] + F->p[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  for (int i = 0; i < 4; i++) r->l[i] = t[i];
}
static inline void fe_neg(fe* r, const fe* a, const field_t* F) {
  fe z = {{0, 0, 0, 0}};
  fe_sub(r, &z, a, F);
}
static inline void fe_mul(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[i] * b->l[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
    t[i + 4] = (uint64_t)c;
  }
  uint64_t carry2 = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t k = t[i] * F->inv;
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)k * F->p[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
    u128 s = (u128)t[i + 4] + (uint64_t)c + carry2;
    t[i + 4] = (uint64_t)s;
    carry2 = (uint64_t)(s >> 64);
  }
  fe_cond_sub(r, t + 4, F);
}
static inline void fe_sqr(fe* r, const fe* a, const field_t* F) { fe_mul(r, a, a, F); }
static void fe_pow(fe* r, const fe* a, const uint64_t e[4], const field_t* F) {
  fe acc = F->r;
  for (int i = 3; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      fe_sqr(&acc, &acc, F);
      if ((e[i] >> b) & 1) fe_mul(&acc, &acc, a, F);
    }
  *r = acc;
}
static void fe_inv(fe* r, const fe* a, const field_t* F) { /* a^(p-2); 0 -> 0 */
  uint64_t e[4] = {F->p[0] - 2, F->p[1], F->p[2], F->p[3]};
  fe_pow(r, a, e, F);
}
static inline void fe_from_mont(uint64_t out[4], const fe* a, const field_t* F) { /* to_repr */
  fe one = {{1, 0, 0, 0}}, t;
  fe_mul(&t, a, &one, F);
  memcpy(out, t.l, 32);
}

/* ---- G1 (halo2curves new_curve_impl!: Jacobian add-2007-bl / dbl-2009-l / madd-2007-bl) ---- */
static inline int j_is_id(const g1j* p) { return fe_is_zero(&p->z); }
static inline void j_set_id(g1j* p) { memset(p, 0, sizeof *p); p->y = FQ.r; }
static void j_double(g1j* r, const g1j* p) {
  if (j_is_id(p)) { *r = *p; return; }
  const field_t* F = &FQ;
  fe a, b, c, d, e, f, t, x3, y3, z3;
  fe_sqr(&a, &p->x, F);
  fe_sqr(&b, &p->y, F);
  fe_sqr(&c, &b, F);
  fe_add(&d, &p->x, &b, F); fe_sqr(&d, &d, F); fe_sub(&d, &d, &a, F); fe_sub(&d, &d, &c, F); fe_add(&d, &d, &d, F);
  fe_add(&e, &a, &a, F); fe_add(&e, &e, &a, F);
  fe_sqr(&f, &e, F);
  fe_mul(&z3, &p->z, &p->y, F); fe_add(&z3, &z3, &z3, F);
  fe_sub(&x3, &f, &d, F); fe_sub(&x3, &x3, &d, F);
  fe_add(&c, &c, &c, F); fe_add(&c, &c, &c, F); fe_add(&c, &c, &c, F);
  fe_sub(&t, &d, &x3, F); fe_mul(&y3, &e, &t, F); fe_sub(&y3, &y3, &c, F);
  r->x = x3; r->y = y3; r->z = z3;
}
static void j_add(g1j* r, const g1j* p, const g1j* q) {
  if (j_is_id(p)) { *r = *q; return; }
  if (j_is_id(q)) { *r = *p; return; }
  const field_t* F = &FQ;
  fe z1z1, z2z2, u1, u2, s1, s2, t;
  fe_sqr(&z1z1, &p->z, F);
  fe_sqr(&z2z2, &q->z, F);
  fe_mul(&u1, &p->x, &z2z2, F);
  fe_mul(&u2, &q->x, &z1z1, F);
  fe_mul(&t, &q->z, &z2z2, F); fe_mul(&s1, &p->y, &t, F);
  fe_mul(&t, &p->z, &z1z1, F); fe_mul(&s2, &q->y, &t, F);
  if (fe_eq(&u1, &u2)) {
    if (fe_eq(&s1, &s2)) { j_double(r, p); } else { j_set_id(r); }
    return;
  }
  fe h, i, j, rr, v, x3, y3, z3;
  fe_sub(&h, &u2, &u1, F);
  fe_add(&i, &h, &h, F); fe_sqr(&i, &i, F);
  fe_mul(&j, &h, &i, F);
  fe_sub(&rr, &s2, &s1, F); fe_add(&rr, &rr, &rr, F);
  fe_mul(&v, &u1, &i, F);
  fe_sqr(&x3, &rr, F); fe_sub(&x3, &x3, &j, F); fe_sub(&x3, &x3, &v, F); fe_sub(&x3, &x3, &v, F);
  fe_mul(&s1, &s1, &j, F); fe_add(&s1, &s1, &s1, F);
  fe_sub(&t, &v, &x3, F); fe_mul(&y3, &rr, &t, F); fe_sub(&y3, &y3, &s1, F);
  fe_add(&z3, &p->z, &q->z, F); fe_sqr(&z3, &z3, F); fe_sub(&z3, &z3, &z1z1, F); fe_sub(&z3, &z3, &z2z2, F);
  fe_mul(&z3, &z3, &h, F);
  r->x = x3; r->y = y3; r->z = z3;
}
static inline int a_is_id(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static void j_from_affine(g1j* r, const g1a* p) {
  if (a_is_id(p)) { j_set_id(r); return; }
  r->x = p->x; r->y = p->y; r->z = FQ.r;
}
static void j_add_affine(g1j* r, const g1j* p, const g1a* q) {
  if (a_is_id(q)) { *r = *p; return; }
  if (j_is_id(p)) { j_from_affine(r, q); return; }
  const field_t* F = &FQ;
  fe z1z1, u2, s2, t;
  fe_sqr(&z1z1, &p->z, F);
  fe_mul(&u2, &q->x, &z1z1, F);
  fe_mul(&t, &p->z, &z1z1, F); fe_mul(&s2, &q->y, &t, F);
  if (fe_eq(&p->x, &u2)) {
    if (fe_eq(&p->y, &s2)) { j_double(r, p); } else { j_set_id(r); }
    return;
  }
  fe h, hh, i, j, rr, v, x3, y3, z3;
  fe_sub(&h, &u2, &p->x, F);
  fe_sqr(&hh, &h, F);
  fe_add(&i, &hh, &hh, F); fe_add(&i, &i, &i, F);
  fe_mul(&j, &h, &i, F);
  fe_sub(&rr, &s2, &p->y, F); fe_add(&rr, &rr, &rr, F);
  fe_mul(&v, &p->x, &i, F);
  fe_sqr(&x3, &rr, F); fe_sub(&x3, &x3, &j, F); fe_sub(&x3, &x3, &v, F); fe_sub(&x3, &x3, &v, F);
  fe_mul(&j, &p->y, &j, F); fe_add(&j, &j, &j, F);
  fe_sub(&t, &v, &x3, F); fe_mul(&y3, &rr, &t, F); fe_sub(&y3, &y3, &j, F);
  fe_add(&z3, &p->z, &h, F); fe_sqr(&z3, &z3, F); fe_sub(&z3, &z3, &z1z1, F); fe_sub(&z3, &z3, &hh, F);
  r->x = x3; r->y = y3; r->z = z3;
}
static void j_to_affine(g1a* r, const g1j* p) {
  if (j_is_id(p)) { memset(r, 0, sizeof *r); return; }
  fe zi, zi2, zi3;
  fe_inv(&zi, &p->z, &FQ);
  fe_sqr(&zi2, &zi, &FQ);
  fe_mul(&zi3, &zi2, &zi, &FQ);
  fe_mul(&r->x, &p->x, &zi2, &FQ);
  fe_mul(&r->y, &p->y, &zi3, &FQ);
}

/* ---- best_multiexp (halo2_proofs arithmetic.rs) ---- */
typedef struct { int kind; g1a a; g1j p; } bucket_t; /* 0 None, 1 Affine, 2 Projective */

static size_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) {
  size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
  if (skip_bytes >= 32) return 0;
  uint8_t v[8] = {0};
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
  uint64_t tmp;
  memcpy(&tmp, v, 8);
  tmp >>= skip_bits - skip_bytes * 8;
  tmp %= (1ull << c);
  return (size_t)tmp;
}

static void multiexp_serial(const fe* coeffs, const g1a* bases, size_t n, g1j* acc) {
  uint8_t(*repr)[32] = malloc(n * 32);
  for (size_t i = 0; i < n; i++) fe_from_mont((uint64_t*)repr[i], &coeffs[i], &FR);
  size_t c;
  if (n < 4) c = 1; else if (n < 32) c = 3; else c = (size_t)ceil(log((double)n));
  size_t segments = 256 / c + 1;
  size_t nb = ((size_t)1 << c) - 1;
  bucket_t* buckets = malloc(nb * sizeof(bucket_t));
  for (size_t seg = segments; seg-- > 0;) {
    for (size_t i = 0; i < c; i++) j_double(acc, acc);
    for (size_t b = 0; b < nb; b++) buckets[b].kind = 0;
    for (size_t i = 0; i < n; i++) {
      size_t co = get_at(seg, c, repr[i]);
      if (co == 0) continue;
      bucket_t* b = &buckets[co - 1];
      if (b->kind == 0) { b->kind = 1; b->a = bases[i]; }
      else if (b->kind == 1) { g1j t; j_from_affine(&t, &b->a); j_add_affine(&b->p, &t, &bases[i]); b->kind = 2; }
      else { j_add_affine(&b->p, &b->p, &bases[i]); }
    }
    g1j running; j_set_id(&running);
    for (size_t b = nb; b-- > 0;) {
      if (buckets[b].kind == 1) j_add_affine(&running, &running, &buckets[b].a);
      else if (buckets[b].kind == 2) j_add(&running, &running, &buckets[b].p);
      j_add(acc, acc, &running);
    }
  }
  free(buckets);
  free(repr);
}

void zgo_best_multiexp(const fe* coeffs, const g1a* bases, size_t n, int threads, g1j* out) {
  if (threads < 1) threads = omp_get_max_threads();
  j_set_id(out);
  if (n > (size_t)threads) {
    size_t chunk = n / threads;
    size_t num_chunks = (n + chunk - 1) / chunk;
    g1j* results = malloc(num_chunks * sizeof(g1j));
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (size_t ci = 0; ci < num_chunks; ci++) {
      size_t b = ci * chunk, e = b + chunk < n ? b + chunk : n;
      j_set_id(&results[ci]);
      multiexp_serial(coeffs + b, bases + b, e - b, &results[ci]);
    }
    for (size_t ci = 0; ci < num_chunks; ci++) j_add(out, out, &results[ci]);
    free(results);
  } else {
    multiexp_serial(coeffs, bases, n, out);
  }
}

void zgo_g1_to_affine(const g1j* in, g1a* out, size_t n) {
  for (size_t i = 0; i < n; i++) j_to_affine(&out[i], &in[i]);
}
/* naive reference: sum of double-and-add scalar multiples (small n only) */
void zgo_msm_naive(const fe* coeffs, const g1a* bases, size_t n, g1j* out) {
  j_set_id(out);
  for (size_t i = 0; i < n; i++) {
    uint64_t s[4];
    fe_from_mont(s, &coeffs[i], &FR);
    g1j acc; j_set_id(&acc);
    for (int w = 3; w >= 0; w--)
      for (int b = 63; b >= 0; b--) {
        j_double(&acc, &acc);
        if ((s[w] >> b) & 1) j_add_affine(&acc, &acc, &bases[i]);
      }
    j_add(out, out, &acc);
  }
}
/* [s^i]G for i < n from a known toy secret (test SRS); g_lagrange via zgo_srs_lagrange */
void zgo_srs_monomial(const fe* s, const g1a* gen, size_t n, g1a* out) {
  fe cur = FR.r;
  fe* pw = malloc(n * sizeof(fe));
  for (size_t i = 0; i < n; i++) { pw[i] = cur; fe_mul(&cur, &cur, s, &FR); }
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < n; i++) {
    uint64_t e[4];
    fe_from_mont(e, &pw[i], &FR);
    g1j acc; j_set_id(&acc);
    for (int w = 3; w >= 0; w--)
      for (int b = 63; b >= 0; b--) {
        j_double(&acc, &acc);
        if ((e[w] >> b) & 1) j_add_affine(&acc, &acc, gen);
      }
    j_to_affine(&out[i], &acc);
  }
  free(pw);
}
/* scalar multiples of one generator by arbitrary Fr scalars (used for g_lagrange = [L_i(s)]G) */
void zgo_g1_mul_many(const fe* scalars, const g1a* gen, size_t n, g1a* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < n; i++) {
    uint64_t e[4];
    fe_from_mont(e, &scalars[i], &FR);
    g1j acc; j_set_id(&acc);
    for (int w = 3; w >= 0; w--)
      for (int b = 63; b >= 0; b--) {
        j_double(&acc, &acc);
        if ((e[w] >> b) & 1) j_add_affine(&acc, &acc, gen);
      }
    j_to_affine(&out[i], &acc);
  }
}

/* ---- best_fft (halo2_proofs arithmetic.rs, v2023_04_20: bit-reverse, serial twiddle scan,
 * recursive_butterfly_arithmetic with rayon::join -> OpenMP tasks) ---- */
static size_t bitreverse(size_t n, size_t l) {
  size_t r = 0;
  for (size_t i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
  return r;
}
static void butterflies(fe* a, size_t n, size_t twiddle_chunk, const fe* tw) {
  if (n == 2) {
    fe t = a[1];
    a[1] = a[0];
    fe_add(&a[0], &a[0], &t, &FR);
    fe_sub(&a[1], &a[1], &t, &FR);
    return;
  }
  fe* left = a; fe* right = a + n / 2;
  if (n >= 4096) {
#pragma omp task
    butterflies(left, n / 2, twiddle_chunk * 2, tw);
#pragma omp task
    butterflies(right, n / 2, twiddle_chunk * 2, tw);
#pragma omp taskwait
  } else {
    butterflies(left, n / 2, twiddle_chunk * 2, tw);
    butterflies(right, n / 2, twiddle_chunk * 2, tw);
  }
  fe t = right[0];
  right[0] = left[0];
  fe_add(&left[0], &left[0], &t, &FR);
  fe_sub(&right[0], &right[0], &t, &FR);
  for (size_t i = 1; i < n / 2; i++) {
    fe_mul(&t, &right[i], &tw[i * twiddle_chunk], &FR);
    right[i] = left[i];
    fe_add(&left[i], &left[i], &t, &FR);
    fe_sub(&right[i], &right[i], &t, &FR);
  }
}
void zgo_best_fft(fe* a, const fe* omega, uint32_t log_n, int threads) {
  if (threads < 1) threads = omp_get_max_threads();
  size_t n = (size_t)1 << log_n;
  for (size_t k = 0; k < n; k++) {
    size_t rk = bitreverse(k, log_n);
    if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
  }
  if (n == 1) return;
  fe* tw = malloc((n / 2) * sizeof(fe));
  fe w = FR.r;
  for (size_t i = 0; i < n / 2; i++) { tw[i] = w; fe_mul(&w, &w, omega, &FR); }
#pragma omp parallel num_threads(threads)
#pragma omp single
  butterflies(a, n, 1, tw);
  free(tw);
}

/* elementwise helpers used by the Python-side restatement of EvaluationDomain / the prover */
void zgo_fr_mul_vec(const fe* a, const fe* b, fe* o, size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) fe_mul(&o[i], &a[i], &b[i], &FR);
}
void zgo_fr_add_vec(const fe* a, const fe* b, fe* o, size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) fe_add(&o[i], &a[i], &b[i], &FR);
}
void zgo_fr_sub_vec(const fe* a, const fe* b, fe* o, size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) fe_sub(&o[i], &a[i], &b[i], &FR);
}
void zgo_fr_scale_vec(const fe* a, const fe* s, fe* o, size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) fe_mul(&o[i], &a[i], s, &FR);
}
/* a[i] *= pw[i % 3]  (distribute_powers_zeta) */
void zgo_fr_scale_mod3(fe* a, const fe pw[3], size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) if (i % 3) fe_mul(&a[i], &a[i], &pw[i % 3], &FR);
}
/* Montgomery batch inversion, zeros stay zero (ff::BatchInvert) */
void zgo_fr_batch_invert(fe* a, size_t n) {
  fe* pre = malloc(n * sizeof(fe));
  fe acc = FR.r;
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    if (!fe_is_zero(&a[i])) fe_mul(&acc, &acc, &a[i], &FR);
  }
  fe inv;
  fe_inv(&inv, &acc, &FR);
  for (size_t i = n; i-- > 0;) {
    if (fe_is_zero(&a[i])) continue;
    fe t;
    fe_mul(&t, &inv, &pre[i], &FR);
    fe_mul(&inv, &inv, &a[i], &FR);
    a[i] = t;
  }
  free(pre);
}
/* eval_polynomial: Horner */
void zgo_fr_eval_poly(const fe* c, size_t n, const fe* x, fe* out) {
  fe acc = {{0, 0, 0, 0}};
  for (size_t i = n; i-- > 0;) {
    fe_mul(&acc, &acc, x, &FR);
    fe_add(&acc, &acc, &c[i], &FR);
  }
  *out = acc;
}
/* to / from canonical little-endian integers */
void zgo_fr_from_mont_vec(const fe* a, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) fe_from_mont(out + 4 * i, &a[i], &FR);
}
void zgo_fr_to_mont_vec(const uint64_t* in, fe* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fe t;
    memcpy(t.l, in + 4 * i, 32);
    fe_mul(&out[i], &t, &FR.r2, &FR);
  }
}
void zgo_fq_from_mont_vec(const fe* a, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) fe_from_mont(out + 4 * i, &a[i], &FQ);
}
void zgo_fq_to_mont_vec(const uint64_t* in, fe* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fe t;
    memcpy(t.l, in + 4 * i, 32);
    fe_mul(&out[i], &t, &FQ.r2, &FQ);
  }
}
/* serial running product as lookup::prover::commit_product / permutation::prover::commit:
 * z[0] = start, z[i] = z[i-1] * f[i-1] for i < n_out */
void zgo_fr_running_product(const fe* f, const fe* start, fe* z, size_t n_out) {
  z[0] = *start;
  for (size_t i = 1; i < n_out; i++) fe_mul(&z[i], &z[i - 1], &f[i - 1], &FR);
}
/* arithmetic::kate_division: q(X) = (a(X) - a(z)) / (X - z), q has n-1 coefficients */
void zgo_fr_kate_division(const fe* a, size_t n, const fe* z, fe* q) {
  fe tmp = {{0, 0, 0, 0}}, negz;
  fe_neg(&negz, z, &FR);
  for (size_t i = n - 1; i >= 1; i--) {
    fe lead;
    fe_sub(&lead, &a[i], &tmp, &FR);
    q[i - 1] = lead;
    fe_mul(&tmp, &lead, &negz, &FR);
  }
}
/* out[i] = a[i] * s + b[i]   (polynomial folding: acc * theta + expr, acc * xn + piece, ...) */
void zgo_fr_mul_add_scalar(const fe* a, const fe* s, const fe* b, fe* o, size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) {
    fe t;
    fe_mul(&t, &a[i], s, &FR);
    fe_add(&o[i], &t, &b[i], &FR);
  }
}
/* Fr::from_u512 on n x 8 little-endian u64 words (halo2curves Fr::random draws) -> Montgomery */
void zgo_fr_from_u512(const uint64_t* w, fe* out, size_t n) {
  fe r3;
  fe_mul(&r3, &FR.r2, &FR.r2, &FR); /* R^2 * R^2 / R = R^3 */
  for (size_t i = 0; i < n; i++) {
    fe d0, d1, t0, t1;
    memcpy(d0.l, w + 8 * i, 32);
    memcpy(d1.l, w + 8 * i + 4, 32);
    fe_mul(&t0, &d0, &FR.r2, &FR);
    fe_mul(&t1, &d1, &r3, &FR);
    fe_add(&out[i], &t0, &t1, &FR);
  }
}
/* rand_xorshift::XorShiftRng + Fr::random (eight next_u64 -> from_u512), `count` draws */
void zgo_xorshift_fr(uint32_t st[4], size_t count, fe* out) {
  uint32_t x = st[0], y = st[1], z = st[2], w = st[3];
  fe r3;
  fe_mul(&r3, &FR.r2, &FR.r2, &FR);
  for (size_t i = 0; i < count; i++) {
    uint64_t wd[8];
    for (int j = 0; j < 8; j++) {
      uint32_t t = x ^ (x << 11);
      x = y; y = z; z = w;
      w = w ^ (w >> 19) ^ (t ^ (t >> 8));
      uint64_t lo = w;
      t = x ^ (x << 11);
      x = y; y = z; z = w;
      w = w ^ (w >> 19) ^ (t ^ (t >> 8));
      wd[j] = lo | ((uint64_t)w << 32);
    }
    fe d0, d1, t0, t1;
    memcpy(d0.l, wd, 32);
    memcpy(d1.l, wd + 4, 32);
    fe_mul(&t0, &d0, &FR.r2, &FR);
    fe_mul(&t1, &d1, &r3, &FR);
    fe_add(&out[i], &t0, &t1, &FR);
  }
  st[0] = x; st[1] = y; st[2] = z; st[3] = w;
}
/* out[i] = start * w^i */
void zgo_fr_powers(const fe* w, const fe* start, fe* out, size_t n) {
  fe cur = *start;
  for (size_t i = 0; i < n; i++) { out[i] = cur; fe_mul(&cur, &cur, w, &FR); }
}
/* lookup::prover::permute_expression_pair on the first `usable` rows (Montgomery in / out):
 * sort the inputs by canonical value; first occurrences copy their value into the table column and
 * consume one table copy; the remaining table values, ascending, fill the repeated rows from the
 * HIGHEST repeated row down (upstream: BTreeMap iteration + repeated_input_rows.pop()).
 * returns 0, or -1 when an input value is missing from the table (Error::ConstraintSystemFailure). */
static int cmp_canon(const void* a, const void* b) {
  const uint64_t* x = (const uint64_t*)a; const uint64_t* y = (const uint64_t*)b;
  for (int i = 3; i >= 0; i--) { if (x[i] < y[i]) return -1; if (x[i] > y[i]) return 1; }
  return 0;
}
int zgo_permute_expression_pair(const fe* a, const fe* s, size_t usable, fe* pa, fe* ps) {
  uint64_t(*A)[4] = malloc(usable * 32);
  uint64_t(*T)[4] = malloc(usable * 32);
  for (size_t i = 0; i < usable; i++) { fe_from_mont(A[i], &a[i], &FR); fe_from_mont(T[i], &s[i], &FR); }
  qsort(A, usable, 32, cmp_canon);
  qsort(T, usable, 32, cmp_canon);
  /* multiset of the table as (unique value, remaining count) in ascending order */
  size_t nu = 0;
  size_t* ustart = malloc((usable + 1) * sizeof(size_t));
  uint32_t* left = malloc(usable * sizeof(uint32_t));
  for (size_t i = 0; i < usable; i++)
    if (i == 0 || cmp_canon(T[i], T[i - 1]) != 0) { ustart[nu] = i; left[nu] = 0; nu++; }
  ustart[nu] = usable;
  for (size_t u = 0; u < nu; u++) left[u] = (uint32_t)(ustart[u + 1] - ustart[u]);
  uint64_t(*S)[4] = calloc(usable, 32);
  size_t* repeated = malloc(usable * sizeof(size_t));
  size_t nrep = 0;
  int rc = 0;
  for (size_t row = 0; row < usable && rc == 0; row++) {
    if (row == 0 || cmp_canon(A[row], A[row - 1]) != 0) {
      memcpy(S[row], A[row], 32);
      size_t lo = 0, hi = nu;
      while (lo < hi) { size_t mid = (lo + hi) / 2; if (cmp_canon(T[ustart[mid]], A[row]) < 0) lo = mid + 1; else hi = mid; }
      if (lo == nu || cmp_canon(T[ustart[lo]], A[row]) != 0 || left[lo] == 0) rc = -1; else left[lo]--;
    } else {
      repeated[nrep++] = row;
    }
  }
  if (rc == 0) {
    for (size_t u = 0; u < nu; u++)
      for (uint32_t c = 0; c < left[u]; c++) memcpy(S[repeated[--nrep]], T[ustart[u]], 32);
    for (size_t i = 0; i < usable; i++) {
      fe t;
      memcpy(t.l, A[i], 32); fe_mul(&pa[i], &t, &FR.r2, &FR);
      memcpy(t.l, S[i], 32); fe_mul(&ps[i], &t, &FR.r2, &FR);
    }
  }
  free(A); free(T); free(S); free(ustart); free(left); free(repeated);
  return rc;
}
/* fixed-base scalar multiplication of one generator by many scalars: 8-bit windows, 32 tables of 255
 * multiples, mixed additions, one shared inversion per 1024 outputs.  Test-SRS generation only. */
void zgo_g1_fixed_base_mul_many(const fe* scalars, const g1a* gen, size_t n, g1a* out) {
  static g1a table[32][255];
  g1j base; j_from_affine(&base, gen);
  for (int w = 0; w < 32; w++) {
    g1j acc = base;
    g1j row[255];
    for (int i = 0; i < 255; i++) { row[i] = acc; j_add(&acc, &acc, &base); }
    for (int i = 0; i < 255; i++) j_to_affine(&table[w][i], &row[i]);
    base = acc; /* 256 * base */
  }
  const size_t CH = 1024;
  size_t nch = (n + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < nch; c++) {
    size_t b = c * CH, e = b + CH < n ? b + CH : n, m = e - b;
    g1j* jac = malloc(m * sizeof(g1j));
    fe* pre = malloc(m * sizeof(fe));
    for (size_t i = 0; i < m; i++) {
      uint8_t bytes[32];
      fe_from_mont((uint64_t*)bytes, &scalars[b + i], &FR);
      g1j acc; j_set_id(&acc);
      for (int w = 0; w < 32; w++) if (bytes[w]) j_add_affine(&acc, &acc, &table[w][bytes[w] - 1]);
      jac[i] = acc;
    }
    fe run = FQ.r;
    for (size_t i = 0; i < m; i++) { pre[i] = run; if (!j_is_id(&jac[i])) fe_mul(&run, &run, &jac[i].z, &FQ); }
    fe inv; fe_inv(&inv, &run, &FQ);
    for (size_t i = m; i-- > 0;) {
      if (j_is_id(&jac[i])) { memset(&out[b + i], 0, sizeof(g1a)); continue; }
      fe zi, zi2, zi3;
      fe_mul(&zi, &inv, &pre[i], &FQ);
      fe_mul(&inv, &inv, &jac[i].z, &FQ);
      fe_sqr(&zi2, &zi, &FQ);
      fe_mul(&zi3, &zi2, &zi, &FQ);
      fe_mul(&out[b + i].x, &jac[i].x, &zi2, &FQ);
      fe_mul(&out[b + i].y, &jac[i].y, &zi3, &FQ);
    }
    free(jac); free(pre);
  }
}
int zgo_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for all host cores explicitly */
void zgo_set_num_threads(int t) { if (t > 0) omp_set_num_threads(t); }

/* Synthetic SRS-shaped bases for benchmarks: out[i] = [start + i + 1] * gen, affine.
 * (A real ParamsKZG basis is [s^i]G; for timing an MSM any distinct curve points do.)
 * Chunks are seeded by one scalar multiplication each and walked with mixed additions;
 * the Jacobian points are normalised with one shared inversion per chunk. */
void zgo_g1_sequence(const g1a* gen, size_t n, g1a* out) {
  const size_t CH = 4096;
  size_t nch = (n + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < nch; c++) {
    size_t b = c * CH, e = b + CH < n ? b + CH : n, m = e - b;
    g1j* jac = malloc(m * sizeof(g1j));
    fe* pre = malloc(m * sizeof(fe));
    g1j acc; j_set_id(&acc);
    uint64_t k = b + 1;
    for (int bit = 63; bit >= 0; bit--) {
      j_double(&acc, &acc);
      if ((k >> bit) & 1) j_add_affine(&acc, &acc, gen);
    }
    for (size_t i = 0; i < m; i++) {
      jac[i] = acc;
      j_add_affine(&acc, &acc, gen);
    }
    fe run = FQ.r;
    for (size_t i = 0; i < m; i++) { pre[i] = run; fe_mul(&run, &run, &jac[i].z, &FQ); }
    fe inv; fe_inv(&inv, &run, &FQ);
    for (size_t i = m; i-- > 0;) {
      fe zi, zi2, zi3;
      fe_mul(&zi, &inv, &pre[i], &FQ);
      fe_mul(&inv, &inv, &jac[i].z, &FQ);
      fe_sqr(&zi2, &zi, &FQ);
      fe_mul(&zi3, &zi2, &zi, &FQ);
      fe_mul(&out[b + i].x, &jac[i].x, &zi2, &FQ);
      fe_mul(&out[b + i].y, &jac[i].y, &zi3, &FQ);
    }
    free(jac); free(pre);
  }
}

/* ======================================================================================================
 * BN254 optimal-ate pairing, for the verifier's final check e(W', [s]G2) = e(R, G2) (halo2_proofs
 * poly/kzg/strategy.rs SingleStrategy::finalize -> DualMSM::check -> Bn256::multi_miller_loop + final_exponentiation,
 * reached from /root/reference/src/wnn.rs:271-280).  Written for clarity, not speed: G2 arithmetic in affine
 * coordinates on the sextic twist E'(Fq2): y^2 = x^3 + 3/(9+i); Fq12 as Fq[w]/(w^12 - 18 w^6 + 82) (w^6 = 9 + i) with
 * schoolbook products; the final exponentiation is one square-and-multiply with the exponent (p^12 - 1)/r supplied
 * by the caller.  Pinned by tests/test_oracle.py: G2 generator on the twist and of order r, bilinearity
 * e(aP, bQ) = e(P, Q)^(ab), non-degeneracy, e(P, Q)^r = 1.
 * ====================================================================================================== */
typedef struct { fe a, b; } f2;          /* a + b i, i^2 = -1 */
typedef struct { f2 x, y; } g2a;         /* affine point on the twist; identity = (0, 0) */
typedef struct { fe c[12]; } f12;        /* sum c[k] w^k */

static fe fq_small(uint64_t v) {         /* v -> Montgomery */
  fe t = {{v, 0, 0, 0}}, r;
  fe_mul(&r, &t, &FQ.r2, &FQ);
  return r;
}
static inline void f2_add(f2* r, const f2* x, const f2* y) { fe_add(&r->a, &x->a, &y->a, &FQ); fe_add(&r->b, &x->b, &y->b, &FQ); }
static inline void f2_sub(f2* r, const f2* x, const f2* y) { fe_sub(&r->a, &x->a, &y->a, &FQ); fe_sub(&r->b, &x->b, &y->b, &FQ); }
static inline void f2_neg(f2* r, const f2* x) { fe_neg(&r->a, &x->a, &FQ); fe_neg(&r->b, &x->b, &FQ); }
static inline void f2_conj(f2* r, const f2* x) { r->a = x->a; fe_neg(&r->b, &x->b, &FQ); }
static inline int f2_is_zero(const f2* x) { return fe_is_zero(&x->a) && fe_is_zero(&x->b); }
static inline int f2_eq(const f2* x, const f2* y) { return fe_eq(&x->a, &y->a) && fe_eq(&x->b, &y->b); }
static void f2_mul(f2* r, const f2* x, const f2* y) {
  fe ac, bd, ad, bc;
  fe_mul(&ac, &x->a, &y->a, &FQ); fe_mul(&bd, &x->b, &y->b, &FQ);
  fe_mul(&ad, &x->a, &y->b, &FQ); fe_mul(&bc, &x->b, &y->a, &FQ);
  fe_sub(&r->a, &ac, &bd, &FQ);
  fe_add(&r->b, &ad, &bc, &FQ);
}
static void f2_mul_fe(f2* r, const f2* x, const fe* s) { fe_mul(&r->a, &x->a, s, &FQ); fe_mul(&r->b, &x->b, s, &FQ); }
static void f2_inv(f2* r, const f2* x) {   /* conj(x) / (a^2 + b^2) */
  fe aa, bb, n, ni;
  fe_mul(&aa, &x->a, &x->a, &FQ); fe_mul(&bb, &x->b, &x->b, &FQ);
  fe_add(&n, &aa, &bb, &FQ);
  fe_inv(&ni, &n, &FQ);
  fe_mul(&r->a, &x->a, &ni, &FQ);
  fe nb;
  fe_neg(&nb, &x->b, &FQ);
  fe_mul(&r->b, &nb, &ni, &FQ);
}
static void f2_pow(f2* r, const f2* x, const uint64_t e[4]) {
  f2 acc = {FQ.r, {{0, 0, 0, 0}}};
  for (int i = 3; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      f2_mul(&acc, &acc, &acc);
      if ((e[i] >> b) & 1) f2_mul(&acc, &acc, x);
    }
  *r = acc;
}
static f2 f2_xi(void) { f2 x = {fq_small(9), FQ.r}; return x; }   /* 9 + i */
static f2 g2_b(void) {                                            /* 3 / (9 + i) */
  f2 xi = f2_xi(), inv, three = {fq_small(3), {{0, 0, 0, 0}}}, r;
  f2_inv(&inv, &xi);
  f2_mul(&r, &three, &inv);
  return r;
}

static inline int g2_is_id(const g2a* p) { return f2_is_zero(&p->x) && f2_is_zero(&p->y); }
/* slope of the chord / tangent through p and q on the twist; returns 0 when the line is vertical */
static int g2_slope(f2* m, const g2a* p, const g2a* q) {
  f2 num, den, inv;
  if (f2_eq(&p->x, &q->x)) {
    if (!f2_eq(&p->y, &q->y) || f2_is_zero(&p->y)) return 0;
    f2 xx, three = {fq_small(3), {{0, 0, 0, 0}}};
    f2_mul(&xx, &p->x, &p->x);
    f2_mul(&num, &xx, &three);
    f2_add(&den, &p->y, &p->y);
  } else {
    f2_sub(&num, &q->y, &p->y);
    f2_sub(&den, &q->x, &p->x);
  }
  f2_inv(&inv, &den);
  f2_mul(m, &num, &inv);
  return 1;
}
static void g2_add(g2a* r, const g2a* p, const g2a* q) {
  if (g2_is_id(p)) { *r = *q; return; }
  if (g2_is_id(q)) { *r = *p; return; }
  f2 m;
  if (!g2_slope(&m, p, q)) { memset(r, 0, sizeof *r); return; }
  f2 mm, x3, t, y3;
  f2_mul(&mm, &m, &m);
  f2_sub(&x3, &mm, &p->x);
  f2_sub(&x3, &x3, &q->x);
  f2_sub(&t, &p->x, &x3);
  f2_mul(&y3, &m, &t);
  f2_sub(&y3, &y3, &p->y);
  r->x = x3;
  r->y = y3;
}
static void g2_mul_bits(g2a* r, const g2a* p, const uint64_t e[4]) {
  g2a acc;
  memset(&acc, 0, sizeof acc);
  for (int i = 3; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      g2_add(&acc, &acc, &acc);
      if ((e[i] >> b) & 1) g2_add(&acc, &acc, p);
    }
  *r = acc;
}

/* ---- Fq12 ---- */
static void f12_one(f12* r) { memset(r, 0, sizeof *r); r->c[0] = FQ.r; }
static int f12_is_one(const f12* x) {
  if (!fe_eq(&x->c[0], &FQ.r)) return 0;
  for (int k = 1; k < 12; k++) if (!fe_is_zero(&x->c[k])) return 0;
  return 1;
}
static void f12_mul(f12* r, const f12* x, const f12* y) {
  static fe c18, c82;
  static int init = 0;
  if (!init) { c18 = fq_small(18); c82 = fq_small(82); init = 1; }
  fe t[23], p;
  memset(t, 0, sizeof t);
  for (int i = 0; i < 12; i++) {
    if (fe_is_zero(&x->c[i])) continue;
    for (int j = 0; j < 12; j++) {
      fe_mul(&p, &x->c[i], &y->c[j], &FQ);
      fe_add(&t[i + j], &t[i + j], &p, &FQ);
    }
  }
  for (int k = 22; k >= 12; k--) {       /* w^k = 18 w^(k-6) - 82 w^(k-12) */
    fe_mul(&p, &t[k], &c18, &FQ);
    fe_add(&t[k - 6], &t[k - 6], &p, &FQ);
    fe_mul(&p, &t[k], &c82, &FQ);
    fe_sub(&t[k - 12], &t[k - 12], &p, &FQ);
  }
  memcpy(r->c, t, sizeof r->c);
}
/* embed s * w^k for s = a + b i in Fq2: i = w^6 - 9  =>  s w^k = (a - 9 b) w^k + b w^(k+6)   (k < 6) */
static void f12_add_f2_wk(f12* r, const f2* s, int k) {
  fe nine = fq_small(9), t;
  fe_mul(&t, &s->b, &nine, &FQ);
  fe_sub(&t, &s->a, &t, &FQ);
  fe_add(&r->c[k], &r->c[k], &t, &FQ);
  fe_add(&r->c[k + 6], &r->c[k + 6], &s->b, &FQ);
}
/* line through the twisted points t1, t2 (tangent if equal), evaluated at P = (xp, yp) in G1:
 * with X = x' w^2, Y = y' w^3 and slope m' w:   l(P) = -yp + (m' xp) w + (y1' - m' x1') w^3 */
static void line_eval(f12* l, const g2a* t1, const g2a* t2, const g1a* P) {
  memset(l, 0, sizeof *l);
  f2 m;
  if (!g2_slope(&m, t1, t2)) {           /* vertical: xp - x1' w^2 */
    l->c[0] = P->x;
    f2 nx;
    f2_neg(&nx, &t1->x);
    f12_add_f2_wk(l, &nx, 2);
    return;
  }
  fe_neg(&l->c[0], &P->y, &FQ);
  f2 c1, c3, mx;
  f2_mul_fe(&c1, &m, &P->x);
  f12_add_f2_wk(l, &c1, 1);
  f2_mul(&mx, &m, &t1->x);
  f2_sub(&c3, &t1->y, &mx);
  f12_add_f2_wk(l, &c3, 3);
}
static void miller_loop(f12* f, const g1a* P, const g2a* Q) {
  f12_one(f);
  if (a_is_id(P) || g2_is_id(Q)) return;
  /* 6u + 2 = 29793968203157093288 (65 bits), u = 4965661367192848881 */
  const uint64_t lo = 0x9d797039be763ba8ull;   /* low 64 bits; bit 64 is the implicit leading one */
  g2a R = *Q;
  f12 l;
  for (int i = 63; i >= 0; i--) {
    f12_mul(f, f, f);
    line_eval(&l, &R, &R, P);
    f12_mul(f, f, &l);
    g2_add(&R, &R, &R);
    if ((lo >> i) & 1) {
      line_eval(&l, &R, Q, P);
      f12_mul(f, f, &l);
      g2_add(&R, &R, Q);
    }
  }
  /* Q1 = pi_p(Q), Q2 = pi_p^2(Q) on the twist: (conj(x') g2, conj(y') g3), g2 = xi^((p-1)/3), g3 = xi^((p-1)/2) */
  static f2 g2c, g3c;
  static fe n2, n3;
  static int init = 0;
  if (!init) {
    uint64_t pm1[4], e3[4], e2[4];
    memcpy(pm1, FQ.p, 32);
    pm1[0] -= 1;
    /* (p-1)/2 */
    for (int k = 0; k < 4; k++) e2[k] = (pm1[k] >> 1) | (k < 3 ? pm1[k + 1] << 63 : 0);
    /* (p-1)/3 by long division */
    u128 rem = 0;
    for (int k = 3; k >= 0; k--) { u128 cur = (rem << 64) | pm1[k]; e3[k] = (uint64_t)(cur / 3); rem = cur % 3; }
    f2 xi = f2_xi();
    f2_pow(&g2c, &xi, e3);
    f2_pow(&g3c, &xi, e2);
    /* xi^((p^2-1)/3) = g2 * conj(g2) = norm(g2) in Fq, likewise for g3 */
    fe aa, bb;
    fe_mul(&aa, &g2c.a, &g2c.a, &FQ); fe_mul(&bb, &g2c.b, &g2c.b, &FQ); fe_add(&n2, &aa, &bb, &FQ);
    fe_mul(&aa, &g3c.a, &g3c.a, &FQ); fe_mul(&bb, &g3c.b, &g3c.b, &FQ); fe_add(&n3, &aa, &bb, &FQ);
    init = 1;
  }
  g2a Q1, nQ2;
  f2 cx, cy;
  f2_conj(&cx, &Q->x); f2_conj(&cy, &Q->y);
  f2_mul(&Q1.x, &cx, &g2c);
  f2_mul(&Q1.y, &cy, &g3c);
  f2_mul_fe(&nQ2.x, &Q->x, &n2);
  f2_mul_fe(&nQ2.y, &Q->y, &n3);
  f2_neg(&nQ2.y, &nQ2.y);
  line_eval(&l, &R, &Q1, P);
  f12_mul(f, f, &l);
  g2_add(&R, &R, &Q1);
  line_eval(&l, &R, &nQ2, P);
  f12_mul(f, f, &l);
}
static void f12_pow_words(f12* r, const f12* x, const uint64_t* e, size_t nwords) {
  f12 acc;
  f12_one(&acc);
  int started = 0;
  for (size_t i = nwords; i-- > 0;)
    for (int b = 63; b >= 0; b--) {
      if (started) f12_mul(&acc, &acc, &acc);
      if ((e[i] >> b) & 1) { f12_mul(&acc, &acc, x); started = 1; }
    }
  *r = acc;
}

/* G2 generator of alt_bn128 (EIP-197), canonical little-endian limbs: x.c0, x.c1, y.c0, y.c1 */
static const uint64_t G2_GEN_CANON[16] = {
  0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull,
  0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull,
  0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull,
  0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull};

void zgo_g2_generator(g2a* out) {
  const fe* w = (const fe*)G2_GEN_CANON;
  fe_mul(&out->x.a, &w[0], &FQ.r2, &FQ); fe_mul(&out->x.b, &w[1], &FQ.r2, &FQ);
  fe_mul(&out->y.a, &w[2], &FQ.r2, &FQ); fe_mul(&out->y.b, &w[3], &FQ.r2, &FQ);
}
int zgo_g2_on_curve(const g2a* p) {
  if (g2_is_id(p)) return 1;
  f2 yy, xx, xxx, b = g2_b(), rhs;
  f2_mul(&yy, &p->y, &p->y);
  f2_mul(&xx, &p->x, &p->x);
  f2_mul(&xxx, &xx, &p->x);
  f2_add(&rhs, &xxx, &b);
  return f2_eq(&yy, &rhs);
}
/* out = [scalar] p, scalar an Fr element in Montgomery form */
void zgo_g2_mul(const g2a* p, const fe* scalar, g2a* out) {
  uint64_t e[4];
  fe_from_mont(e, scalar, &FR);
  g2_mul_bits(out, p, e);
}
/* f = final_exponentiation(miller(P, Q)); exp = (p^12 - 1) / r as little-endian 64-bit words */
void zgo_pairing(const g1a* P, const g2a* Q, const uint64_t* exp, size_t nwords, f12* out) {
  f12 f;
  miller_loop(&f, P, Q);
  f12_pow_words(out, &f, exp, nwords);
}
/* 1 iff prod_i e(P_i, Q_i) == 1 */
int zgo_pairing_check(const g1a* P, const g2a* Q, size_t n, const uint64_t* exp, size_t nwords) {
  f12 acc, f;
  f12_one(&acc);
  for (size_t i = 0; i < n; i++) {
    miller_loop(&f, &P[i], &Q[i]);
    f12_mul(&acc, &acc, &f);
  }
  f12_pow_words(&f, &acc, exp, nwords);
  return f12_is_one(&f);
}
void zgo_f12_mul(const f12* a, const f12* b, f12* out) { f12_mul(out, a, b); }
void zgo_f12_pow(const f12* a, const uint64_t* exp, size_t nwords, f12* out) { f12_pow_words(out, a, exp, nwords); }
