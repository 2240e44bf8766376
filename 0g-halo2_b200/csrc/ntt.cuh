// NTT pass descriptors shared by ntt.cu and the context layer.
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace zg {

#ifndef ZG_NTT_THREADS
#define ZG_NTT_THREADS 128
#endif
#ifndef ZG_NTT_LOGC
#define ZG_NTT_LOGC 2
#endif
constexpr int NTT_THREADS = ZG_NTT_THREADS;
#ifndef ZG_NTT_MAX_S
#define ZG_NTT_MAX_S 8
#endif
constexpr uint32_t NTT_MAX_S = ZG_NTT_MAX_S;  // index bits per pass (tile = 2^S rows)
constexpr uint32_t NTT_LOGC = ZG_NTT_LOGC;   // columns per tile (4 x 32 B = one 128-B line)

enum : uint32_t {
  NTT_IN_COSET = 1,   // multiply input i by in_scale[i % 3]   (distribute_powers_zeta)
  NTT_OUT_SCALE = 2,  // multiply output i by out_scale[0]      (1/n of the inverse transform)
  NTT_OUT_MOD3 = 4,   // ... by out_scale[i % 3] instead         (1/n folded with zeta^-i)
};

struct NttPassArgs {
  const Fr* in;
  Fr* out;
  size_t in_stride, out_stride;  // elements between consecutive polynomials of a batch
  const Fr* tw;                  // stage-major twiddles
  uint32_t logn, lo, S, logc, contiguous;
  uint32_t n_in;   // inputs at index >= n_in read as zero (zero-padding of coeff_to_extended)
  uint32_t n_out;  // outputs at index >= n_out are not stored (truncation of extended_to_coeff)
  uint32_t flags;
  Fr in_scale[3];
  Fr out_scale[3];
  // per-element scale tables and several cosets per polynomial (see NttPlan)
  uint32_t cosets;
  // coset slot z of this launch is coset id base + z * step for the input offset / the output offset / the tables
  // (a rank that owns a subset of the cosets; dense slots in the scratch passes in between)
  uint32_t in_cid_base, in_cid_step, out_cid_base, out_cid_step, tab_cid_base, tab_cid_step;
  size_t in_coset_stride, out_coset_stride, in_table_stride, out_table_stride;
  const Fr* in_table;
  const Fr* out_table;
};

struct NttPlan {
  const Fr* in;
  Fr* out;
  Fr* tmp;  // scratch of batch * cosets * tmp_stride elements, needed when logn > NTT_MAX_S
  size_t in_stride, out_stride, tmp_stride;
  const Fr* tw;
  uint32_t logn, batch;
  uint32_t n_in, n_out, flags;
  Fr in_scale[3];
  Fr out_scale[3];
  // `cosets` transforms per polynomial that differ in the per-element input table (coset c multiplies input i by
  // in_table[c * in_table_stride + i]) or output table: the prover's 3 x 2n extended domain (extdomain.cuh)
  uint32_t cosets = 1;
  uint32_t coset_first = 0, coset_step = 1;     // the launch covers coset ids coset_first + z * coset_step, z < cosets
  size_t in_coset_stride = 0, out_coset_stride = 0;
  const Fr* in_table = nullptr;
  size_t in_table_stride = 0;
  const Fr* out_table = nullptr;
  size_t out_table_stride = 0;
};

cudaError_t ntt_build_twiddles(Fr* tab, Fr* scratch_flat, const Fr& w, uint32_t logn,
                               cudaStream_t stream);
cudaError_t ntt_run(const NttPlan& plan, cudaStream_t stream, uint64_t* launch_counter);

}  // namespace zg
