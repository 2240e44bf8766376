// ChaCha20 keystream RNG keyed from the operating system: the production RNG of the backend (the role
// rand::rngs::OsRng plays at /root/reference/src/wnn.rs:256; include/zg_b200.h declares the entry points).
// Host code, plain C++ (no CUDA): RFC 8439 block function with a 64-bit block counter in words 12-13 and the nonce in
// 14-15.  A k = 17 proof draws 8.4 MB (the random polynomial of the vanishing argument), so whole runs of 8 blocks are
// generated lane-parallel -- the loops over `l` vectorise (AVX2 clone picked at load time, baseline SSE2 otherwise).
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <sys/random.h>
#include "../../include/zg_b200.h"

namespace {

constexpr int LANES = 8;

typedef uint32_t v8u __attribute__((vector_size(32)));   // 8 blocks side by side (GCC vector extension: AVX2 or 2 x SSE2)

#define ZG_ROTV(v, c) (((v) << (c)) | ((v) >> (32 - (c))))
#define ZG_QRV(a, b, c, d)                                   \
  a += b; d ^= a; d = ZG_ROTV(d, 16); c += d; b ^= c; b = ZG_ROTV(b, 12); \
  a += b; d ^= a; d = ZG_ROTV(d, 8);  c += d; b ^= c; b = ZG_ROTV(b, 7);

// LANES consecutive blocks starting at `counter` -> out[LANES * 64]
__attribute__((target_clones("avx2", "default")))
void chacha20_blocks(const uint32_t key[8], uint64_t counter, const uint32_t nonce[2], uint8_t* out) {
  static const uint32_t sigma[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  v8u in[16], x[16];
  for (int i = 0; i < 4; i++) in[i] = (v8u){} + sigma[i];
  for (int i = 0; i < 8; i++) in[4 + i] = (v8u){} + key[i];
  for (int l = 0; l < LANES; l++) {
    const uint64_t c = counter + (uint64_t)l;
    in[12][l] = (uint32_t)c;
    in[13][l] = (uint32_t)(c >> 32);
  }
  in[14] = (v8u){} + nonce[0];
  in[15] = (v8u){} + nonce[1];
  for (int i = 0; i < 16; i++) x[i] = in[i];
  for (int round = 0; round < 10; round++) {
    ZG_QRV(x[0], x[4], x[8], x[12]) ZG_QRV(x[1], x[5], x[9], x[13]) ZG_QRV(x[2], x[6], x[10], x[14]) ZG_QRV(x[3], x[7], x[11], x[15])
    ZG_QRV(x[0], x[5], x[10], x[15]) ZG_QRV(x[1], x[6], x[11], x[12]) ZG_QRV(x[2], x[7], x[8], x[13]) ZG_QRV(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; i++) x[i] += in[i];
  // word i of block l is x[i][l]: two 8 x 8 transposes put every block's 64 bytes together (little-endian hosts only)
  for (int half = 0; half < 2; half++) {
    v8u* r = x + 8 * half;
    v8u a[8], b[8], c[8];
    for (int p = 0; p < 4; p++) {
      a[2 * p] = __builtin_shuffle(r[2 * p], r[2 * p + 1], (v8u){0, 8, 1, 9, 4, 12, 5, 13});
      a[2 * p + 1] = __builtin_shuffle(r[2 * p], r[2 * p + 1], (v8u){2, 10, 3, 11, 6, 14, 7, 15});
    }
    for (int q = 0; q < 2; q++) {
      b[4 * q + 0] = __builtin_shuffle(a[4 * q + 0], a[4 * q + 2], (v8u){0, 1, 8, 9, 4, 5, 12, 13});
      b[4 * q + 1] = __builtin_shuffle(a[4 * q + 0], a[4 * q + 2], (v8u){2, 3, 10, 11, 6, 7, 14, 15});
      b[4 * q + 2] = __builtin_shuffle(a[4 * q + 1], a[4 * q + 3], (v8u){0, 1, 8, 9, 4, 5, 12, 13});
      b[4 * q + 3] = __builtin_shuffle(a[4 * q + 1], a[4 * q + 3], (v8u){2, 3, 10, 11, 6, 7, 14, 15});
    }
    for (int p = 0; p < 4; p++) {
      c[p] = __builtin_shuffle(b[p], b[p + 4], (v8u){0, 1, 2, 3, 8, 9, 10, 11});
      c[p + 4] = __builtin_shuffle(b[p], b[p + 4], (v8u){4, 5, 6, 7, 12, 13, 14, 15});
    }
    for (int l = 0; l < LANES; l++) memcpy(out + 64 * l + 32 * half, &c[l], 32);
  }
}
#undef ZG_QRV
#undef ZG_ROTV

}  // namespace

extern "C" {

void zg_chacha20_seed(zg_chacha20* r, const uint8_t key[32]) {
  memset(r, 0, sizeof(*r));
  memcpy(r->key, key, 32);
}

int zg_chacha20_seed_os(zg_chacha20* r) {
  uint8_t key[32];
  size_t got = 0;
  while (got < sizeof(key)) {
    ssize_t k = getrandom(key + got, sizeof(key) - got, 0);
    if (k < 0) {
      if (errno == EINTR) continue;
      return ZG_E_STATE;
    }
    got += (size_t)k;
  }
  zg_chacha20_seed(r, key);
  memset(key, 0, sizeof(key));
  return ZG_OK;
}

void zg_chacha20_fill(void* state, uint64_t* out, size_t n) {
  zg_chacha20* r = (zg_chacha20*)state;
  uint8_t* dst = (uint8_t*)out;
  size_t left = n * 8;
  uint8_t tmp[LANES * 64];
  while (left) {
    if (r->have) {                                   // bytes left over from the previous call come first
      const size_t take = left < r->have ? left : r->have;
      memcpy(dst, r->buf + (64 - r->have), take);
      r->have -= (uint32_t)take; dst += take; left -= take;
      continue;
    }
    if (left >= ((size_t)1 << 20)) {
      // a k = 17 proof draws 8.4 MB for its random polynomial while the GPU commits the advice columns (2.3 ms): the
      // keystream is seekable, so a large request is cut into ranges of whole block runs generated side by side
      const size_t runs = left / sizeof(tmp);
      static const int T = [] {                     // two threads by default: several lanes and ranks share the host
        const char* e = getenv("ZG_RNG_THREADS");
        int v = e ? atoi(e) : 2;
        return v < 1 ? 1 : v > 8 ? 8 : v;
      }();
      const size_t per = (runs + T - 1) / T;
      std::thread th[8];
      for (int t = 0; t < T; t++) {
        const size_t r0 = (size_t)t * per, r1 = r0 + per < runs ? r0 + per : runs;
        th[t] = std::thread([=]() {
          for (size_t q = r0; q < r1; q++) chacha20_blocks(r->key, r->counter + q * LANES, r->nonce, dst + q * sizeof(tmp));
        });
      }
      for (int t = 0; t < T; t++) th[t].join();
      r->counter += runs * LANES;
      dst += runs * sizeof(tmp); left -= runs * sizeof(tmp);
      continue;
    }
    if (left >= sizeof(tmp)) {                       // whole runs of LANES blocks go straight to the destination
      chacha20_blocks(r->key, r->counter, r->nonce, dst);
      r->counter += LANES;
      dst += sizeof(tmp); left -= sizeof(tmp);
      continue;
    }
    // tail: generate LANES blocks, hand out what is needed; keep at most one partial block, rewind the counter over
    // the blocks that were not touched so that the stream does not depend on how it is cut into calls
    chacha20_blocks(r->key, r->counter, r->nonce, tmp);
    const size_t whole = left / 64, rem = left % 64;
    memcpy(dst, tmp, left);
    r->counter += whole + (rem ? 1 : 0);
    if (rem) {
      memcpy(r->buf, tmp + 64 * whole, 64);
      r->have = (uint32_t)(64 - rem);
    }
    left = 0;
  }
}

}  // extern "C"
