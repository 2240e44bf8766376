// CPU-side harness for the experimental FP64 Montgomery multiplication (csrc/experimental/field52.cuh): the rounding
// mode is set to toward-zero so that std::fma reproduces the device's __fma_rz.  Test infrastructure only.
#include <cfenv>
#include "../../0g-halo2_b200/csrc/experimental/field52.cuh"
using namespace zg52;
extern "C" {
// a, b, out: n x 4 u64 words; p: 4 words; pinv = -p^-1 mod 2^52
void h_mul52(const uint64_t* a, const uint64_t* b, uint64_t* out, int n, const uint64_t* p, uint64_t pinv) {
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  Params52 P;
  Fe52 pl = from_words(p);
  for (int k = 0; k < 5; k++) P.p[k] = pl.l[k];
  P.pinv = pinv;
  for (int i = 0; i < n; i++) {
    Fe52 r = mul(from_words(a + 4 * i), from_words(b + 4 * i), P);
    to_words(r, out + 4 * i);
  }
  fesetround(old);
}
}
