// Integer-pipe micro-benchmarks: the denominators of the MSM / NTT / quotient rooflines
// (SURVEY.md section 8d asks for a measured, dependent-free IMAD peak next to every result).
#include "ctx.cuh"

using namespace zg;

namespace {

constexpr int MB_THREADS = 256;

__global__ void __launch_bounds__(MB_THREADS) mb_imad_kernel(uint32_t* out, uint32_t b, uint32_t c, uint32_t iters) {
  uint32_t a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x + i;
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = a[i] * b + c;
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= a[i];
  if (s == 0x12345678u) out[0] = s;
}

__global__ void __launch_bounds__(MB_THREADS) mb_imad_wide_kernel(uint64_t* out, uint32_t b, uint32_t iters) {
  uint64_t a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = (uint64_t)(uint32_t)(a[i] >> 7) * b + a[i];
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i];
  if (s == 0x12345678u) out[0] = s;
}

// IMAD.WIDE without any other instruction in the loop: 16 independent 64-bit accumulators, multiplier = own low word
__global__ void __launch_bounds__(MB_THREADS) mb_imad_wide16_kernel(uint64_t* out, uint32_t b, uint32_t iters) {
  uint64_t a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x + i;
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (uint64_t)(uint32_t)a[i] * b + a[i];
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= a[i];
  if (s == 0x12345678u) out[0] = s;
}
// FP64 pipe (idle on this path today): DFMA alone, and DFMA interleaved 1:1 with IMAD.WIDE to see whether the two
// pipes issue side by side (the question behind Emmart-style double-precision limb products)
template <bool WITH_IMAD>
__global__ void __launch_bounds__(MB_THREADS) mb_dfma_kernel(double* out, double b, double c, uint32_t ib, uint32_t iters) {
  double d[8];
  uint64_t a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d[i] = 1.0 + threadIdx.x * 1e-9 + i;
    a[i] = threadIdx.x + i;
  }
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      d[i] = fma(d[i], b, c);
      if (WITH_IMAD) a[i] = (uint64_t)(uint32_t)a[i] * ib + a[i];
    }
  }
  double s = 0;
  uint64_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    s += d[i];
    x ^= a[i];
  }
  if (s == 0.123456789 || x == 0x12345678u) out[0] = s + (double)x;
}

template <int VARIANT>
__global__ void __launch_bounds__(MB_THREADS) mb_mulmod_kernel(Fr* out, Fr seed, uint32_t iters) {
  Fr x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    x[i] = seed;
    x[i].v[0] ^= (threadIdx.x + 64 * i) & 0xff;
  }
  Fr y = seed;
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
#if defined(__CUDA_ARCH__)
      if (VARIANT == 0) x[i] = fp_mul_portable(x[i], y);
      else if (VARIANT == 1) x[i] = fp_mul_ptx(x[i], y);
      else if (VARIANT == 2) x[i] = fp_mul_eo(x[i], y);
      else if (VARIANT == 3) x[i] = fp_sqr_fast(x[i]);
      else x[i] = fp_mul2add(x[i], y, y, x[i]);   // counted as two products
#endif
    }
  }
  Fr s = fp_add(fp_add(x[0], x[1]), fp_add(x[2], x[3]));
  if (s.v[7] == 0xffffffffu) out[0] = s;   // never true (values < r), keeps the chain alive
  if (blockIdx.x == 0 && threadIdx.x == 0) out[1] = s;
}

}  // namespace

extern "C" int zg_bench_int_pipe(zg_ctx* ctx, int kind, uint32_t iters, double* giga_per_s) {
  ZG_ENTER(ctx);
  if (!giga_per_s || iters == 0) return ctx->fail(ZG_E_INVALID, "bench_int_pipe: bad arguments");
  cudaDeviceProp prop;
  ZG_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
  const int blocks = prop.multiProcessorCount * 8;
  void* scratch = nullptr;
  ZG_CUDA(cudaMalloc(&scratch, 4096));
  cudaEvent_t e0, e1;
  ZG_CUDA(cudaEventCreate(&e0));
  ZG_CUDA(cudaEventCreate(&e1));
  Fr seed = host_fr_root_of_unity();
  double ops_per_thread = 0;
  for (int rep = 0; rep < 2; rep++) {  // rep 0 = warm-up
    ZG_CUDA(cudaEventRecord(e0, ctx->stream));
    switch (kind) {
      case 0:
        mb_imad_kernel<<<blocks, MB_THREADS, 0, ctx->stream>>>((uint32_t*)scratch, 0x9e3779b1u, 12345u, iters);
        ops_per_thread = 16.0 * iters;
        break;
      case 1:
        mb_imad_wide_kernel<<<blocks, MB_THREADS, 0, ctx->stream>>>((uint64_t*)scratch, 0x9e3779b1u, iters);
        ops_per_thread = 8.0 * iters;
        break;
      case 2:
        mb_mulmod_kernel<0><<<blocks, MB_THREADS, 0, ctx->stream>>>((Fr*)scratch, seed, iters);
        ops_per_thread = 4.0 * iters;
        break;
      case 3:
        mb_mulmod_kernel<1><<<blocks, MB_THREADS, 0, ctx->stream>>>((Fr*)scratch, seed, iters);
        ops_per_thread = 4.0 * iters;
        break;
      case 4:
        mb_mulmod_kernel<2><<<blocks, MB_THREADS, 0, ctx->stream>>>((Fr*)scratch, seed, iters);
        ops_per_thread = 4.0 * iters;
        break;
      case 8:   // dedicated squaring
        mb_mulmod_kernel<3><<<blocks, MB_THREADS, 0, ctx->stream>>>((Fr*)scratch, seed, iters);
        ops_per_thread = 4.0 * iters;
        break;
      case 9:   // two-product multiply-add with one reduction: rate in PRODUCTS per second
        mb_mulmod_kernel<4><<<blocks, MB_THREADS, 0, ctx->stream>>>((Fr*)scratch, seed, iters);
        ops_per_thread = 8.0 * iters;
        break;
      case 5:
        mb_dfma_kernel<false><<<blocks, MB_THREADS, 0, ctx->stream>>>((double*)scratch, 0.999999, 1e-7, 0x9e3779b1u, iters);
        ops_per_thread = 8.0 * iters;
        break;
      case 6:   // reported rate = DFMA + IMAD.WIDE instructions together
        mb_dfma_kernel<true><<<blocks, MB_THREADS, 0, ctx->stream>>>((double*)scratch, 0.999999, 1e-7, 0x9e3779b1u, iters);
        ops_per_thread = 16.0 * iters;
        break;
      case 7:
        mb_imad_wide16_kernel<<<blocks, MB_THREADS, 0, ctx->stream>>>((uint64_t*)scratch, 0x9e3779b1u, iters);
        ops_per_thread = 16.0 * iters;
        break;
      default:
        return ctx->fail(ZG_E_INVALID, "bench_int_pipe: unknown kind");
    }
    ctx->launches++;
    ZG_CUDA(cudaEventRecord(e1, ctx->stream));
    ZG_CUDA(cudaEventSynchronize(e1));
  }
  float ms = 0;
  ZG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *giga_per_s = ops_per_thread * (double)blocks * MB_THREADS / (ms * 1e-3) / 1e9;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ZG_CUDA(cudaFree(scratch));
  return ZG_OK;
}

// ---- diagnostics: element-wise field ops on the device (used by the -m gpu parity tests) ----
namespace {
template <class P>
__global__ void dbg_field_kernel(int op, const Fp<P>* a, const Fp<P>* b, Fp<P>* o, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#if defined(__CUDA_ARCH__)
  Fp<P> x = a[i], y = b[i], r;
  switch (op) {
    case 0: r = fp_mul(x, y); break;
    case 1: r = fp_mul_portable(x, y); break;
    case 2: r = fp_mul_ptx(x, y); break;
    case 3: r = fp_add(x, y); break;
    case 4: r = fp_sub(x, y); break;
    case 5: r = fp_inv(x); break;
    case 6: r = fp_from_mont(x); break;
    case 7: r = fp_to_mont(x); break;
    case 8: r = fp_mul_eo(x, y); break;
    case 9: r = fp_sqr_fast(x); break;                                   // dedicated squaring (field_gen.cuh)
    case 10: r = fp_mul2add(x, y, fp_add(x, y), x); break;               // x*y + (x+y)*x, one reduction
    case 11: r = fp_mul2add(x, x, fp_neg_lazy(y), y); break;             // x^2 - y^2 (y = 0: the operand is p itself)
    default: r = fp_zero<P>(); break;
  }
  o[i] = r;
#endif
}
}  // namespace

extern "C" int zg_debug_field_op(zg_ctx* ctx, int field, int op, const void* a, const void* b, void* out,
                                 size_t n) {
  ZG_ENTER(ctx);
  if (n == 0) return ZG_OK;
  if (op < 0 || op > 11 || field < 0 || field > 1) return ctx->fail(ZG_E_INVALID, "debug_field_op: bad op/field");
  int rc = ws_reserve(ctx, ctx->ws_stage, 96 * n);
  if (rc) return rc;
  uint8_t* d = ctx->ws_stage.p;
  ZG_CUDA(cudaMemcpyAsync(d, a, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  ZG_CUDA(cudaMemcpyAsync(d + 32 * n, b, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t blocks = (uint32_t)((n + 127) / 128);
  if (field == 0)
    dbg_field_kernel<FrParams><<<blocks, 128, 0, ctx->stream>>>(op, (const Fr*)d, (const Fr*)(d + 32 * n), (Fr*)(d + 64 * n), (uint32_t)n);
  else
    dbg_field_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>(op, (const Fq*)d, (const Fq*)(d + 32 * n), (Fq*)(d + 64 * n), (uint32_t)n);
  ctx->launches++;
  ZG_CUDA(cudaGetLastError());
  ZG_CUDA(cudaMemcpyAsync(out, d + 64 * n, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
  ZG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZG_OK;
}
