// Native witness synthesis of the WNN circuit (host code; nothing here touches the GPU).
//
// Restates the ASSIGNMENT half of zero_g's circuit -- `WnnCircuit::synthesize` -> `WnnChip::predict`
// (/root/reference/src/gadgets/wnn.rs:180-237, 372-393) and the sub-chips it drives: encode_image.rs, greater_than.rs,
// range_check.rs (+ halo2_gadgets' LookupRangeCheckConfig), bits2num.rs, hash.rs, bloom_filter.rs with array_lookup.rs,
// byte_selector.rs, bit_selector.rs, and_bits.rs, response_accumulator.rs -- under halo2's SimpleFloorPlanner: regions
// are placed in call order, each at the first row that is free in every advice column it touches.  The output is the six
// advice columns in the layout `zg_create_proof` takes (Montgomery limbs, rows >= usable left for the blinding scalars).
// In the reference this work is Rust on the host (BASELINE north_star: "WnnChip witness synthesis stays on the host");
// here it is the native counterpart of the Python front-end (zg_b200/plonk/gadgets.py), checked cell for cell against
// it by tests/test_wnn_synth.py, so that a proof service is not throttled by a 0.5 s-per-image interpreter loop.
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "../../include/zg_b200.h"
#include "field.cuh"

using namespace zg;

namespace {

struct U256 {
  uint64_t l[4];
};
inline U256 u256(uint64_t x) { return U256{{x, 0, 0, 0}}; }
inline bool fits128(const U256& a) { return (a.l[2] | a.l[3]) == 0; }
inline unsigned __int128 lo128(const U256& a) { return ((unsigned __int128)a.l[1] << 64) | a.l[0]; }
inline U256 from128(unsigned __int128 x) { return U256{{(uint64_t)x, (uint64_t)(x >> 64), 0, 0}}; }

inline Fr to_mont(const U256& c) {
  Fr raw;
  memcpy(raw.v, c.l, 32);
  return fp_to_mont(raw);
}
inline U256 from_mont(const Fr& m) {
  Fr c = fp_from_mont(m);
  U256 r;
  memcpy(r.l, c.v, 32);
  return r;
}
// (z - word) / 2^shift in the field; an exact integer shift whenever z is a small integer with those low bits
inline U256 sub_shift(const U256& z, uint64_t word, uint32_t shift, const Fr& inv_pow) {
  if (fits128(z)) {
    unsigned __int128 v = lo128(z);
    if (v >= word && (((v - word) & (((unsigned __int128)1 << shift) - 1)) == 0)) return from128((v - word) >> shift);
  }
  return from_mont(fp_mul(fp_sub(to_mont(z), to_mont(u256(word))), inv_pow));
}
inline uint64_t bits_of(const U256& v, uint32_t pos, uint32_t width) {   // (v >> pos) & (2^width - 1), width <= 32
  uint32_t limb = pos >> 6, off = pos & 63;
  uint64_t x = limb < 4 ? v.l[limb] >> off : 0;
  if (off && limb + 1 < 4) x |= v.l[limb + 1] << (64 - off);
  return x & ((1ull << width) - 1);
}

struct Cell {
  uint32_t col, row;
  U256 v;
};

}  // namespace

struct zg_wnn {
  uint64_t p = 0;
  uint32_t l = 0, n_hashes = 0, bits_per_hash = 0, bits_per_filter = 0, n_classes = 0, n_filters = 0;
  uint32_t w = 0, h = 0, nb = 0;
  std::vector<uint16_t> thr;     // [i][j][b], values 0..256
  std::vector<uint64_t> perm;    // input permutation over the w*h*nb thermometer bits
  uint32_t word_index_bits = 0, bytes_per_word = 0, words_per_filter = 0;
  std::vector<uint64_t> words;   // [bloom index][word], big-endian packed bits (array_lookup.rs from_be_bits)
  Fr inv_pow2[33];               // 1 / 2^k, Montgomery
};

// the model is shared read-only by the threads of a proof service: the last error is kept per calling thread
static thread_local std::string g_wnn_err;

namespace {

struct Synth {
  const zg_wnn& W;
  Fr* const* adv;
  uint32_t usable;
  uint32_t height[6] = {0, 0, 0, 0, 0, 0};
  bool overflow = false;
  Fr small[512];   // Montgomery forms of 0..511: almost every assigned value is a bit, a byte or a small index

  Synth(const zg_wnn& w, Fr* const* a, uint32_t u) : W(w), adv(a), usable(u) {
    for (uint32_t i = 0; i < 512; i++) small[i] = fp_from_u64<FrParams>(i);
  }
  // SimpleFloorPlanner: first row free in every touched column; all touched columns then end at start + rows
  uint32_t place(uint32_t colmask, uint32_t rows) {
    uint32_t start = 0;
    for (int c = 0; c < 6; c++)
      if (colmask >> c & 1) start = height[c] > start ? height[c] : start;
    for (int c = 0; c < 6; c++)
      if (colmask >> c & 1) height[c] = start + rows;
    if ((uint64_t)start + rows > usable) overflow = true;   // Error::NotEnoughRowsAvailable
    return start;
  }
  Cell put(uint32_t col, uint32_t row, const U256& v) {
    if (row < usable) {
      if ((v.l[1] | v.l[2] | v.l[3]) == 0 && v.l[0] < 512) adv[col][row] = small[v.l[0]];
      else adv[col][row] = to_mont(v);
    }
    return Cell{col, row, v};
  }
  Cell put(uint32_t col, uint32_t row, uint64_t v) { return put(col, row, u256(v)); }
  void put_fr(uint32_t col, uint32_t row, const Fr& m) {
    if (row < usable) adv[col][row] = m;
  }

  // ---- range_check.rs + halo2_gadgets LookupRangeCheckConfig<F, 8> on advice column 5 -------------------------
  Cell copy_check(const Cell& el, uint32_t num_words) {
    uint32_t s = place(1u << 5, num_words + 1);
    Cell z = put(5, s, el.v);
    for (uint32_t i = 0; i < num_words; i++) {
      uint64_t word = bits_of(el.v, 8 * i, 8);
      z = put(5, s + i + 1, sub_shift(z.v, word, 8, W.inv_pow2[8]));
    }
    return z;
  }
  void copy_short_check(const Cell& el, uint32_t num_bits) {
    uint32_t s = place(1u << 5, 3);
    put(5, s, el.v);
    U256 shifted;
    if (fits128(el.v) && (lo128(el.v) >> 100) == 0) shifted = from128(lo128(el.v) << (8 - num_bits));
    else shifted = from_mont(fp_mul(to_mont(el.v), fp_from_u64<FrParams>(1ull << (8 - num_bits))));
    put(5, s + 1, shifted);
    put_fr(5, s + 2, W.inv_pow2[num_bits]);
  }
  void range_check(const Cell& c, uint32_t n_bits) {
    uint32_t words = n_bits / 8;
    Cell last = c;
    if (words) last = copy_check(c, words);
    if (n_bits % 8) copy_short_check(last, n_bits % 8);
  }
  void le_constant(const Cell& x, uint64_t y) {
    uint32_t s = place(1u << 5, 3);
    put(5, s, x.v);
    put(5, s + 1, y);
    U256 diff;
    if (fits128(x.v) && lo128(x.v) <= y) diff = u256(y - (uint64_t)lo128(x.v));
    else diff = from_mont(fp_sub(to_mont(u256(y)), to_mont(x.v)));
    Cell d = put(5, s + 2, diff);
    uint32_t bl = 0;
    for (uint64_t t = y; t; t >>= 1) bl++;
    range_check(d, bl);
  }

  // ---- greater_than.rs (x = 0, y = 1, diff = 2, is_gt = 3) ----------------------------------------------------
  void gt_row(uint32_t s, const U256& x, uint64_t y, Cell& d, Cell& g) {
    const uint64_t gt = (uint32_t)x.l[0] > y ? 1 : 0;                    // to_u32(x) > y
    const uint64_t rhs = 256 * gt + y;
    U256 diff;
    if (fits128(x) && lo128(x) <= rhs) diff = u256(rhs - (uint64_t)lo128(x));
    else diff = from_mont(fp_sub(to_mont(u256(rhs)), to_mont(x)));
    put(1, s, y);
    d = put(2, s, diff);
    g = put(3, s, gt);
  }
  void gt_witness(uint64_t x, uint64_t y, Cell& xc, Cell& g) {           // greater_than.rs:135-165
    uint32_t s = place(0xF, 1);
    xc = put(0, s, x);
    Cell d;
    gt_row(s, xc.v, y, d, g);
    range_check(xc, 8);
    range_check(g, 1);
    range_check(d, 8);
  }
  Cell gt_copy(const Cell& x, uint64_t y) {                              // :167-191
    uint32_t s = place(0xF, 1);
    put(0, s, x.v);
    Cell d, g;
    gt_row(s, x.v, y, d, g);
    range_check(g, 1);
    range_check(d, 8);
    return g;
  }

  // ---- encode_image.rs:75-150 ------------------------------------------------------------------------------
  void encode(const uint8_t* image, std::vector<Cell>& bits) {
    std::vector<Cell> first((size_t)W.w * W.h);
    std::vector<char> have((size_t)W.w * W.h, 0);
    bits.reserve((size_t)W.w * W.h * W.nb);
    for (uint32_t b = 0; b < W.nb; b++)
      for (uint32_t i = 0; i < W.w; i++)
        for (uint32_t j = 0; j < W.h; j++) {
          const uint32_t t = W.thr[((size_t)i * W.h + j) * W.nb + b];
          const size_t px = (size_t)i * W.h + j;
          Cell cell;
          if (t == 0) {
            uint32_t s = place(1u << 3, 1);                              // "bit is one"
            cell = put(3, s, 1);
          } else if (!have[px]) {
            Cell xc;
            gt_witness(image[px], t - 1, xc, cell);
            first[px] = xc;
            have[px] = 1;
          } else {
            cell = gt_copy(first[px], t - 1);
          }
          bits.push_back(cell);
        }
  }

  // ---- bits2num.rs (input = 3, accumulator = 4) ----------------------------------------------------------------
  Cell bits2num_le(const Cell* const* bits, uint32_t count) {           // bits[0] is the least significant
    uint32_t s = place((1u << 3) | (1u << 4), count + 1);
    put(4, s, 0);
    unsigned __int128 val = 0;
    Cell cell{};
    for (uint32_t i = 0; i < count; i++) {
      const Cell& b = *bits[count - 1 - i];                             // most significant first
      val = val * 2 + (uint64_t)b.v.l[0];
      cell = put(4, s + i + 1, from128(val));
      put(3, s + i, b.v);
    }
    return cell;
  }

  // ---- hash.rs:129-210 (input = 0, quotient = 1, remainder = 2, msb = 3, hash = 4) ----------------------------
  Cell hash(const Cell& inp) {
    uint32_t s = place(0x1F, 1);
    put(0, s, inp.v);
    // cubed = x^3 as an integer (x < 2^64, so < 2^192 < r), q = cubed / p, rem = cubed % p  (utils.rs:47-58)
    const uint64_t x = inp.v.l[0];
    uint64_t c[4] = {0, 0, 0, 0};
    {
      unsigned __int128 sq = (unsigned __int128)x * x;
      unsigned __int128 lo = (unsigned __int128)(uint64_t)sq * x;
      unsigned __int128 hi = (unsigned __int128)(uint64_t)(sq >> 64) * x + (uint64_t)(lo >> 64);
      c[0] = (uint64_t)lo;
      c[1] = (uint64_t)hi;
      c[2] = (uint64_t)(hi >> 64);
    }
    U256 q{{0, 0, 0, 0}};
    unsigned __int128 rem = 0;
    for (int k = 3; k >= 0; k--) {
      unsigned __int128 cur = (rem << 64) | c[k];
      q.l[k] = (uint64_t)(cur / W.p);
      rem = cur % W.p;
    }
    const uint64_t r64 = (uint64_t)rem;
    const uint64_t msb = W.l >= 64 ? 0 : r64 >> W.l;
    const uint64_t hv = W.l >= 64 ? r64 : r64 & ((1ull << W.l) - 1);
    Cell qc = put(1, s, q);
    Cell rc = put(2, s, r64);
    Cell mc = put(3, s, msb);
    Cell out = put(4, s, hv);
    range_check(qc, W.bits_per_filter * 3 - W.l);
    range_check(mc, 1);
    le_constant(rc, W.p - 1);
    return out;
  }

  // ---- bloom_filter/array_lookup.rs:305-456 (hash decomposition = 0, byte index = 1, bit index = 2, bloom index = 3,
  //      bloom value = 4) --------------------------------------------------------------------------------------
  struct Looked { Cell word, byte_index, bit_index; };
  void array_lookup(const Cell& hv, uint64_t bloom_index, std::vector<Looked>& out) {
    const uint32_t nh = W.n_hashes, bph = W.bits_per_hash, nbb = bph - W.word_index_bits;
    uint32_t s = place(0x1F, nh + 1);
    out.resize(nh);
    put(0, s, hv.v);
    U256 dec = hv.v;
    for (uint32_t i = 0; i < nh; i++) {
      const uint64_t h = bits_of(hv.v, bph * i, bph);                  // little-endian hash i
      dec = sub_shift(dec, h, bph, W.inv_pow2[bph]);
      if (i + 1 < nh) put(0, s + i + 1, dec);
      else put(0, s + nh, 0);                                           // assign_advice_from_constant(.., 0)
      const uint32_t h32 = (uint32_t)h;
      const uint64_t wi = h32 >> nbb, by = (h32 & ((1u << nbb) - 1)) >> 3, bi = h32 & 7;
      put(3, s + i, bloom_index);
      out[i].word = put(4, s + i, W.words[bloom_index * W.words_per_filter + wi]);
      out[i].byte_index = put(1, s + i, by);
      out[i].bit_index = put(2, s + i, bi);
    }
  }

  // ---- bloom_filter/byte_selector.rs:184-352 (all six columns) ---------------------------------------------------
  Cell byte_select(const Cell& word, const Cell& index, uint32_t nbytes) {
    uint32_t s = place(0x3F, nbytes + 1);
    const uint32_t idx = (uint32_t)index.v.l[0];
    const uint64_t ith = bits_of(word.v, 8 * (nbytes - 1 - idx), 8);     // bytes_be[idx]
    put(0, s, word.v);
    U256 dec = word.v;
    for (uint32_t i = 0; i < nbytes; i++) {
      dec = sub_shift(dec, bits_of(word.v, 8 * i, 8), 8, W.inv_pow2[8]);
      if (i + 1 < nbytes) put(0, s + i + 1, dec);
      else put(0, s + nbytes, 0);
    }
    for (uint32_t i = 0; i < nbytes; i++) {
      put(1, s + i, index.v);
      put(2, s + nbytes - 1 - i, i);
      put(3, s + i, (nbytes - 1 - i) == idx ? 1 : 0);
    }
    put(4, s, 0);
    for (uint32_t i = 1; i < nbytes; i++) put(4, s + i, (nbytes - i) <= idx ? 1 : 0);
    put(4, s + nbytes, 1);
    Cell result = put(5, s, 0);
    for (uint32_t i = 1; i <= nbytes; i++) result = put(5, s + i, (nbytes - i) <= idx ? ith : 0);
    return result;
  }
  // ---- bloom_filter/bit_selector.rs:138-164 (byte = 0, index = 1, bit = 2) ---------------------------------------
  Cell bit_select(const Cell& byte, const Cell& index) {
    uint32_t s = place(0x7, 1);
    put(0, s, byte.v);
    put(1, s, index.v);
    return put(2, s, (byte.v.l[0] >> (7 - (uint32_t)index.v.l[0])) & 1);
  }
  // ---- bloom_filter/and_bits.rs:83-121 (bits = 4, accumulator = 5) ------------------------------------------------
  Cell and_bits(const std::vector<Cell>& bits) {
    const uint32_t cnt = (uint32_t)bits.size();
    uint32_t s = place((1u << 4) | (1u << 5), cnt + 1);
    Cell cell = put(5, s, 1);
    uint64_t acc = 1;
    for (uint32_t i = 0; i < cnt; i++) {
      put(4, s + i, bits[i].v);
      acc *= bits[i].v.l[0];                                             // bits are 0 / 1
      cell = put(5, s + i + 1, acc);
    }
    return cell;
  }
  // ---- bloom_filter.rs:165-191 --------------------------------------------------------------------------------------
  Cell bloom_lookup(const Cell& hv, uint64_t bloom_index) {
    std::vector<Looked> looked;
    array_lookup(hv, bloom_index, looked);
    std::vector<Cell> bits;
    for (size_t i = looked.size(); i-- > 0;) {                           // the reference iterates the reversed list
      Cell byte = byte_select(looked[i].word, looked[i].byte_index, W.bytes_per_word);
      bits.push_back(bit_select(byte, looked[i].bit_index));
    }
    return and_bits(bits);
  }
  // ---- response_accumulator.rs:77-133 (inputs 0..3, accumulator 4) ----------------------------------------------------
  Cell accumulate(const std::vector<Cell>& resp) {
    const uint32_t nrows = ((uint32_t)resp.size() + 3) / 4;
    uint32_t s = place(0x1F, nrows + 1);
    Cell cell = put(4, s, 0);
    uint64_t acc = 0;
    for (uint32_t row = 0; row < nrows; row++) {
      for (uint32_t i = 0; i < 4; i++) {
        const size_t k = (size_t)row * 4 + i;
        if (k < resp.size()) {
          put(i, s + row, resp[k].v);
          acc += resp[k].v.l[0];
        } else {
          put(i, s + row, 0);
        }
      }
      cell = put(4, s + row + 1, acc);
    }
    return cell;
  }
};

}  // namespace

extern "C" {

int zg_wnn_create(const zg_wnn_desc* d, zg_wnn** out) {
  if (!d || !out) return ZG_E_INVALID;
  *out = nullptr;
  if (!d->thresholds || !d->input_permutation || !d->bloom_bits || d->p == 0 || d->n_hashes == 0 || d->bits_per_hash < 7 ||
      d->bits_per_hash > 32 || d->n_hashes * d->bits_per_hash > 63 || d->bits_per_filter == 0 || d->bits_per_filter > 63)
    return ZG_E_INVALID;
  std::unique_ptr<zg_wnn> w(new zg_wnn());
  w->p = d->p;
  w->n_hashes = d->n_hashes;
  w->bits_per_hash = d->bits_per_hash;
  w->l = d->n_hashes * d->bits_per_hash;
  w->bits_per_filter = d->bits_per_filter;
  w->n_classes = d->n_classes;
  w->n_filters = d->n_filters;
  w->w = d->width;
  w->h = d->height;
  w->nb = d->bits_per_input;
  const size_t nbits = (size_t)w->w * w->h * w->nb;
  if (nbits % w->bits_per_filter || nbits / w->bits_per_filter != w->n_filters) return ZG_E_INVALID;
  w->thr.assign(d->thresholds, d->thresholds + nbits);
  for (uint16_t t : w->thr)
    if (t > 256) return ZG_E_INVALID;
  w->perm.assign(d->input_permutation, d->input_permutation + nbits);
  for (uint64_t pi : w->perm)
    if (pi >= nbits) return ZG_E_INVALID;
  // array_lookup.rs:63-67: byte_index_bits = (bits_per_hash - 3) / 2 - floor(log2(n_hashes)), computed in f64 then truncated
  uint32_t flog = 0;
  while ((2u << flog) <= w->n_hashes) flog++;
  const double bib = (w->bits_per_hash - 3.0) / 2.0 - (double)flog;
  const uint32_t byte_index_bits = bib > 0 ? (uint32_t)bib : 0;
  w->word_index_bits = w->bits_per_hash - (byte_index_bits + 3);
  w->bytes_per_word = 1u << byte_index_bits;
  if (w->bytes_per_word > 8) return ZG_E_INVALID;                        // words are kept in 64 bits
  const uint32_t wl = 8 * w->bytes_per_word;                              // bits per word
  const size_t entries = (size_t)1 << w->bits_per_hash;
  w->words_per_filter = (uint32_t)(entries / wl);
  const size_t nfilters = (size_t)w->n_classes * w->n_filters;
  w->words.assign(nfilters * w->words_per_filter, 0);
  for (size_t f = 0; f < nfilters; f++)
    for (size_t j = 0; j < w->words_per_filter; j++) {
      uint64_t v = 0;
      for (uint32_t t = 0; t < wl; t++) v = (v << 1) | (d->bloom_bits[f * entries + j * wl + t] ? 1u : 0u);   // first bit = MSB
      w->words[f * w->words_per_filter + j] = v;
    }
  Fr half = fp_inv(fp_from_u64<FrParams>(2));
  w->inv_pow2[0] = fp_one<FrParams>();
  for (int k = 1; k <= 32; k++) w->inv_pow2[k] = fp_mul(w->inv_pow2[k - 1], half);
  *out = w.release();
  return ZG_OK;
}

void zg_wnn_free(zg_wnn* w) { delete w; }

const char* zg_wnn_last_error(const zg_wnn* w) { return w ? g_wnn_err.c_str() : "null model"; }

int zg_wnn_synthesize(const zg_wnn* w, const uint8_t* image, uint32_t k, uint32_t usable_rows, zg_fr* const* advice, uint64_t* outputs) {
  if (!w || !image || !advice || k < 1 || k > 28 || usable_rows == 0 || usable_rows > (1u << k)) return ZG_E_INVALID;
  const size_t n = (size_t)1 << k;
  for (int c = 0; c < 6; c++) {
    if (!advice[c]) return ZG_E_INVALID;
    memset((void*)advice[c], 0, n * sizeof(zg_fr));
  }
  Synth S(*w, reinterpret_cast<Fr* const*>(advice), usable_rows);
  std::vector<Cell> bits;
  S.encode(image, bits);
  const uint32_t nb = w->bits_per_filter;
  std::vector<const Cell*> group(nb);
  std::vector<Cell> hashes;
  hashes.reserve(w->n_filters);
  for (uint32_t f = 0; f < w->n_filters; f++) {
    for (uint32_t t = 0; t < nb; t++) group[t] = &bits[w->perm[(size_t)f * nb + t]];
    Cell joint = S.bits2num_le(group.data(), nb);
    hashes.push_back(joint);                                             // replaced by its hash below, in the same order
  }
  for (uint32_t f = 0; f < w->n_filters; f++) hashes[f] = S.hash(hashes[f]);
  // all bloom lookups of every class first, then the accumulations (WnnChip::predict, src/gadgets/wnn.rs:214-236)
  std::vector<std::vector<Cell>> resp(w->n_classes);
  for (uint32_t c = 0; c < w->n_classes; c++) {
    resp[c].reserve(w->n_filters);
    for (uint32_t f = 0; f < w->n_filters; f++) resp[c].push_back(S.bloom_lookup(hashes[f], (uint64_t)c * w->n_filters + f));
  }
  for (uint32_t c = 0; c < w->n_classes; c++) {
    Cell score = S.accumulate(resp[c]);
    if (outputs) outputs[c] = score.v.l[0];
  }
  if (S.overflow) {
    g_wnn_err = "not enough rows available (k = " + std::to_string(k) + ")";   // plonk::Error::NotEnoughRowsAvailable
    return ZG_E_SYNTH;
  }
  return ZG_OK;
}

}  // extern "C"
