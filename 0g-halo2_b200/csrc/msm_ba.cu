// Batched-affine pre-reduction of the sorted (bucket, point) list, ahead of the XYZZ accumulation of msm.cu.
//
// Same role as the first levels of msm_accumulate_kernel (the bucket sums of halo2_proofs' `best_multiexp`, reached from
// /root/reference/src/wnn.rs:242-259), cheaper arithmetic: an affine + affine addition costs 3 products once
// 1 / (x2 - x1) is known, against 10 for the mixed XYZZ addition; Montgomery's trick shares one inversion among
// thousands of additions at 3 more products each.  One ROUND halves every bucket: adjacent entries of a bucket are added
// pairwise (an odd one is carried over), the results form a new sorted list of materialised affine points.
//   k_ba_plan     out_count[b] = ceil(in_count[b] / 2)            (+ exclusive scan -> output offsets)
//   k_ba_forward  thread t owns outputs [16 t, 16 t + 16): product of its pairs' denominators -> totals[t]
//   (inversion)   totals -> 1 / totals, two-level batch inversion (poly.cu, instantiated for Fq)
//   k_ba_backward same walk, prefix products kept per thread, the inverse total peeled backwards into every
//                 1 / (x2 - x1); lambda, x3, y3; writes the point and its (bucket, index) entry
// 7.3 products per addition (1 forward + 2 peel + 1 recomputed prefix + 3 for the sum + the shared inversion).  After two
// rounds a quarter of the entries is left for msm_accumulate_kernel, which reads them like a window table.
// Doubling (equal points meet when a test SRS holds multiples of one generator), P + (-P) and identity operands are
// resolved per pair with a denominator of 2y / 1 / 1.
#define ZG_FP_MUL_NOINLINE 1
#include "msm.cuh"
#include "msm_ba.cuh"

namespace zg {

namespace {

constexpr int BA_M = 16;   // outputs per thread

ZG_HD void ba_load(const BaSrc& s, uint32_t i, Fq& x, Fq& y) {
  const uint2 e = s.ent[i];
  const G1Affine* tb = ((s.alt_mask >> (e.x >> s.log_nb)) & 1u) ? s.alt : s.tab;
  const G1Affine& p = tb[e.y & 0x7fffffffu];
  x = p.x;
  y = p.y;
  if ((e.y >> 31) && !(fp_is_zero(x) && fp_is_zero(y))) y = fp_neg(y);
}

enum : int { BA_COPY1 = 0, BA_COPY2 = 1, BA_ADD = 2, BA_DBL = 3, BA_INF = 4 };

// what the pair (P1, P2) needs and the denominator it contributes to the shared inversion
ZG_HD int ba_classify(const Fq& x1, const Fq& y1, const Fq& x2, const Fq& y2, Fq& d) {
  const bool id1 = fp_is_zero(x1) && fp_is_zero(y1), id2 = fp_is_zero(x2) && fp_is_zero(y2);
  d = fp_one<FqParams>();
  if (id2) return BA_COPY1;
  if (id1) return BA_COPY2;
  if (!fp_eq(x1, x2)) {
    d = fp_sub(x2, x1);
    return BA_ADD;
  }
  if (fp_eq(y1, y2) && !fp_is_zero(y1)) {
    d = fp_dbl(y1);
    return BA_DBL;
  }
  return BA_INF;
}

// bucket of output `o`: largest b with out_off[b] <= o (empty buckets share offsets, so search for the upper bound)
ZG_HD uint32_t ba_find_bucket(const uint32_t* out_off, uint32_t nb, uint32_t o) {
  uint32_t lo = 0, hi = nb;             // invariant: out_off[lo] <= o < out_off[hi]
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (out_off[mid] <= o) lo = mid; else hi = mid;
  }
  return lo;
}

struct BaWalk {
  uint32_t b, o, o_end;
};
ZG_HD bool ba_walk_init(const BaPlanView& p, uint32_t t, BaWalk& w) {
  const uint32_t total = p.out_off[p.nb];
  const uint64_t first = (uint64_t)t * BA_M;
  if (first >= total) return false;
  w.o = (uint32_t)first;
  w.o_end = (first + BA_M < total) ? w.o + BA_M : total;
  w.b = ba_find_bucket(p.out_off, p.nb, w.o);
  return true;
}
// inputs of output w.o: i0 and whether a partner i0 + 1 exists; advances the bucket cursor first
ZG_HD void ba_item(const BaPlanView& p, BaWalk& w, uint32_t& i0, bool& pair) {
  while (w.o >= p.out_off[w.b + 1]) w.b++;
  const uint32_t j = w.o - p.out_off[w.b];
  const uint32_t start = p.in_off[w.b], cnt = p.in_off[w.b + 1] - start;
  i0 = start + 2 * j;
  pair = 2 * j + 1 < cnt;
}

ZG_HD void ba_forward_thread(const BaSrc& s, const BaPlanView& p, uint32_t t, Fq* totals) {
  BaWalk w;
  if (!ba_walk_init(p, t, w)) return;
  Fq acc = fp_one<FqParams>();
  for (; w.o < w.o_end; w.o++) {
    uint32_t i0;
    bool pair;
    ba_item(p, w, i0, pair);
    if (!pair) continue;
    Fq x1, y1, x2, y2, d;
    ba_load(s, i0, x1, y1);
    ba_load(s, i0 + 1, x2, y2);
    ba_classify(x1, y1, x2, y2, d);
    acc = fp_mul(acc, d);
  }
  totals[t] = acc;
}

ZG_HD void ba_backward_thread(const BaSrc& s, const BaPlanView& p, uint32_t t, const Fq* totals_inv, G1Affine* out_pts,
                              uint2* out_ent) {
  BaWalk w;
  if (!ba_walk_init(p, t, w)) return;
  const uint32_t o0 = w.o, n_out = w.o_end - w.o;
  Fq pre[BA_M];
  Fq acc = fp_one<FqParams>();
  // forward again: prefix products (a pair's own denominator is recomputed in the backward loop)
  uint32_t last_b;
  {
    BaWalk f = w;
    for (uint32_t k = 0; k < n_out; k++, f.o++) {
      uint32_t i0;
      bool pair;
      ba_item(p, f, i0, pair);
      pre[k] = acc;
      if (!pair) continue;
      Fq x1, y1, x2, y2, d;
      ba_load(s, i0, x1, y1);
      ba_load(s, i0 + 1, x2, y2);
      ba_classify(x1, y1, x2, y2, d);
      acc = fp_mul(acc, d);
    }
    last_b = f.b;
  }
  Fq inv = totals_inv[t];
  // backward: outputs in decreasing order, the bucket cursor walks down from the bucket of the last output
  uint32_t b = last_b;
  for (uint32_t k = n_out; k-- > 0;) {
    const uint32_t o = o0 + k;
    while (o < p.out_off[b]) b--;
    const uint32_t j = o - p.out_off[b];
    const uint32_t start = p.in_off[b], cnt = p.in_off[b + 1] - start;
    const uint32_t i0 = start + 2 * j;
    Fq x1, y1;
    ba_load(s, i0, x1, y1);
    G1Affine r;
    if (2 * j + 1 >= cnt) {
      r.x = x1;
      r.y = y1;
    } else {
      Fq x2, y2, d;
      ba_load(s, i0 + 1, x2, y2);
      const int kind = ba_classify(x1, y1, x2, y2, d);
      const Fq dinv = fp_mul(inv, pre[k]);      // 1 / d  (1 for the copy / infinity cases, whose d is 1)
      inv = fp_mul(inv, d);
      if (kind == BA_COPY1) {
        r.x = x1; r.y = y1;
      } else if (kind == BA_COPY2) {
        r.x = x2; r.y = y2;
      } else if (kind == BA_INF) {
        r.x = fp_zero<FqParams>(); r.y = fp_zero<FqParams>();
      } else {
        Fq num;
        if (kind == BA_ADD) {
          num = fp_sub(y2, y1);
        } else {                                   // doubling: lambda = 3 x^2 / (2 y)
          Fq xx = fp_sqr(x1);
          num = fp_add(fp_dbl(xx), xx);
        }
        const Fq lam = fp_mul(num, dinv);
        r.x = fp_sub(fp_sub(fp_sqr(lam), x1), x2);
        r.y = fp_sub(fp_mul(lam, fp_sub(x1, r.x)), y1);
      }
    }
    out_pts[o] = r;
    out_ent[o] = make_uint2(b, o);
  }
}

__global__ void k_ba_plan(const uint32_t* __restrict__ in_off, uint32_t nb, uint32_t* __restrict__ out_cnt) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  out_cnt[b] = (in_off[b + 1] - in_off[b] + 1u) >> 1;
}
__global__ void __launch_bounds__(128) k_ba_forward(BaSrc s, BaPlanView p, Fq* __restrict__ totals, uint32_t nthreads) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nthreads) ba_forward_thread(s, p, t, totals);
}
__global__ void __launch_bounds__(128) k_ba_backward(BaSrc s, BaPlanView p, const Fq* __restrict__ totals_inv,
                                                     G1Affine* __restrict__ out_pts, uint2* __restrict__ out_ent, uint32_t nthreads) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nthreads) ba_backward_thread(s, p, t, totals_inv, out_pts, out_ent);
}

}  // namespace

size_t ba_workspace_bytes(uint32_t L_max, uint32_t nb, int rounds) {
  size_t bytes = 0;
  uint64_t len = L_max;
  for (int r = 0; r < rounds; r++) {
    len = (len + nb) / 2 + 1;
    bytes += (size_t)len * (sizeof(G1Affine) + sizeof(uint2)) + 512;   // points + entries of round r
    bytes += ((size_t)nb + 2) * 4 * 2 + 512;                           // counts + offsets
  }
  const size_t tmax = ((size_t)L_max / 2 + nb) / BA_M + 2;
  bytes += tmax * sizeof(Fq) + tmax / 16 * sizeof(Fq) + 1024;          // totals + second-level products
  bytes += scan_scratch_words(nb, 1) * 4 + 256;
  return bytes;
}

// Runs `rounds` halving rounds on the device.  in: sorted entries + offsets (nb + 1, offsets[nb] = L) + table(s).
// out: the source the accumulation kernel should read (entries, their count pointer, point array).
cudaError_t ba_reduce(const BaSrc& src0, const uint32_t* in_off0, uint32_t nb, uint32_t L_max, int rounds, uint8_t* ws,
                      cudaStream_t st, uint64_t* nl, BaSrc* out_src, const uint32_t** out_count_ptr) {
  uint8_t* p = ws;
  auto take = [&](size_t bytes) { uint8_t* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  BaSrc src = src0;
  const uint32_t* in_off = in_off0;
  uint64_t len = L_max;
  const size_t tmax = ((size_t)L_max / 2 + nb) / BA_M + 2;
  Fq* totals = (Fq*)take(tmax * sizeof(Fq));
  Fq* lvl2 = (Fq*)take((tmax / 16 + 64) * sizeof(Fq));
  uint32_t* scan_scratch = (uint32_t*)take(scan_scratch_words(nb, 1) * 4 + 64);
  LaunchCounter lc{nl};
  for (int r = 0; r < rounds; r++) {
    len = (len + nb) / 2 + 1;
    G1Affine* pts = (G1Affine*)take((size_t)len * sizeof(G1Affine));
    uint2* ent = (uint2*)take((size_t)len * sizeof(uint2));
    uint32_t* cnt = (uint32_t*)take(((size_t)nb + 2) * 4);
    uint32_t* off = (uint32_t*)take(((size_t)nb + 2) * 4);
    k_ba_plan<<<(nb + 255) / 256, 256, 0, st>>>(in_off, nb, cnt);
    lc++;
    ScanJobs sj{};
    sj.in[0] = cnt;
    sj.out[0] = off;
    scan_excl_u32(sj, 1, nb, scan_scratch, st, lc);
    BaPlanView pv{in_off, off, nb};
    const uint32_t nthreads = (uint32_t)((len + BA_M - 1) / BA_M);
    k_ba_forward<<<(nthreads + 127) / 128, 128, 0, st>>>(src, pv, totals, nthreads);
    lc++;
    fq_batch_invert(totals, nthreads, st, lc, lvl2);
    k_ba_backward<<<(nthreads + 127) / 128, 128, 0, st>>>(src, pv, totals, pts, ent, nthreads);
    lc++;
    src.ent = ent;
    src.tab = pts;
    src.alt = pts;
    src.alt_mask = 0;
    in_off = off;
  }
  *out_src = src;
  *out_count_ptr = in_off + nb;
  return cudaGetLastError();
}

// ---- host emulation of the same per-thread code (CPU test of the pairing / inversion logic; no GPU needed) -------------
// Builds a table of multiples of the generator, a sorted entry list with heavy buckets, duplicates (doubling), P / -P pairs
// and identity points, runs `rounds` rounds thread by thread on the host and compares every bucket sum with a plain XYZZ
// accumulation.  Returns 0 when they agree.
int ba_host_selftest(uint32_t nb, uint32_t max_per_bucket, int rounds, uint32_t seed) {
  const uint32_t T = 64;
  std::vector<G1Affine> table(T);
  {
    G1Affine g;
    g.x = fp_from_u64<FqParams>(1);
    g.y = fp_from_u64<FqParams>(2);
    G1Xyzz acc = xyzz_identity();
    for (uint32_t i = 0; i < T; i++) {
      if (i == 0) {                               // an identity entry in the table
        table[i].x = fp_zero<FqParams>(); table[i].y = fp_zero<FqParams>();
        continue;
      }
      xyzz_madd(acc, g.x, g.y);
      Fq i3 = fp_inv(acc.zzz);
      Fq tt = fp_mul(acc.zz, i3);
      table[i].x = fp_mul(acc.x, fp_sqr(tt));
      table[i].y = fp_mul(acc.y, i3);
    }
  }
  uint64_t rs = seed * 0x9E3779B97F4A7C15ull + 12345;
  auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); };
  std::vector<uint2> ent;
  std::vector<uint32_t> off(nb + 1);
  for (uint32_t b = 0; b < nb; b++) {
    off[b] = (uint32_t)ent.size();
    uint32_t c = rnd() % (max_per_bucket + 1);
    if (b % 7 == 3) c = 0;                        // empty buckets
    if (b % 11 == 5) c = max_per_bucket * 8;      // a heavy bucket
    for (uint32_t i = 0; i < c; i++) {
      uint32_t idx = rnd() % T, neg = (rnd() & 7) == 0;
      if (b % 5 == 1) idx = 7;                    // all equal: doublings, and with signs P + (-P)
      ent.push_back(make_uint2(b, idx | (neg << 31)));
    }
  }
  off[nb] = (uint32_t)ent.size();
  // reference bucket sums
  std::vector<G1Xyzz> ref(nb, xyzz_identity());
  BaSrc s0{ent.data(), table.data(), table.data(), 0, 31};
  for (uint32_t i = 0; i < ent.size(); i++) {
    Fq x, y;
    ba_load(s0, i, x, y);
    if (!(fp_is_zero(x) && fp_is_zero(y))) xyzz_madd(ref[ent[i].x], x, y);
  }
  BaSrc src = s0;
  std::vector<uint32_t> in_off = off;
  std::vector<std::vector<G1Affine>> keep_pts(rounds);
  std::vector<std::vector<uint2>> keep_ent(rounds);
  std::vector<std::vector<uint32_t>> keep_off(rounds);
  for (int r = 0; r < rounds; r++) {
    std::vector<uint32_t>& o = keep_off[r];
    o.assign(nb + 1, 0);
    for (uint32_t b = 0; b < nb; b++) o[b + 1] = o[b] + ((in_off[b + 1] - in_off[b] + 1) >> 1);
    const uint32_t total = o[nb];
    BaPlanView pv{in_off.data(), o.data(), nb};
    const uint32_t nthreads = (total + BA_M - 1) / BA_M + 1;       // one thread past the end, as a sized-for-the-bound grid has
    std::vector<Fq> totals(nthreads, fp_one<FqParams>());
    for (uint32_t t = 0; t < nthreads; t++) ba_forward_thread(src, pv, t, totals.data());
    for (auto& v : totals) v = fp_inv(v);
    keep_pts[r].assign(total + 1, G1Affine{});
    keep_ent[r].assign(total + 1, make_uint2(0, 0));
    for (uint32_t t = 0; t < nthreads; t++) ba_backward_thread(src, pv, t, totals.data(), keep_pts[r].data(), keep_ent[r].data());
    src.ent = keep_ent[r].data();
    src.tab = src.alt = keep_pts[r].data();
    src.alt_mask = 0;
    in_off = o;
  }
  std::vector<G1Xyzz> got(nb, xyzz_identity());
  for (uint32_t i = 0; i < in_off[nb]; i++) {
    Fq x, y;
    ba_load(src, i, x, y);
    if (src.ent[i].x >= nb) return 2;
    if (!(fp_is_zero(x) && fp_is_zero(y))) xyzz_madd(got[src.ent[i].x], x, y);
  }
  for (uint32_t b = 0; b < nb; b++) {
    // compare projectively: X1 ZZ2 == X2 ZZ1, Y1 ZZZ2 == Y2 ZZZ1, identity == identity
    const bool i1 = xyzz_is_identity(ref[b]), i2 = xyzz_is_identity(got[b]);
    if (i1 != i2) return 3;
    if (i1) continue;
    if (!fp_eq(fp_mul(ref[b].x, got[b].zz), fp_mul(got[b].x, ref[b].zz))) return 4;
    if (!fp_eq(fp_mul(ref[b].y, got[b].zzz), fp_mul(got[b].y, ref[b].zzz))) return 5;
  }
  return 0;
}

}  // namespace zg

extern "C" int zg_debug_ba_selftest(uint32_t nb, uint32_t max_per_bucket, int rounds, uint32_t seed) {
  return zg::ba_host_selftest(nb, max_per_bucket, rounds, seed);
}
