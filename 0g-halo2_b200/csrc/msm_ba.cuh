// Batched-affine pre-reduction of the sorted MSM entry list (msm_ba.cu).
#pragma once
#include <cuda_runtime.h>
#include "curve.cuh"
#include "scan.cuh"

namespace zg {

// where the points of a sorted entry list live: entry i = (bucket key, index | sign << 31) into `tab`, or into `alt`
// when bit (key >> log_nb) of alt_mask is set (a batch that mixes the two SRS bases)
struct BaSrc {
  const uint2* ent;
  const G1Affine* tab;
  const G1Affine* alt;
  uint32_t alt_mask, log_nb;
};
struct BaPlanView {
  const uint32_t* in_off;    // nb + 1: first entry of every bucket in the input list
  const uint32_t* out_off;   // nb + 1: first entry of every bucket in the halved list
  uint32_t nb;
};

size_t ba_workspace_bytes(uint32_t L_max, uint32_t nb, int rounds);
cudaError_t ba_reduce(const BaSrc& src0, const uint32_t* in_off0, uint32_t nb, uint32_t L_max, int rounds, uint8_t* ws,
                      cudaStream_t st, uint64_t* nl, BaSrc* out_src, const uint32_t** out_count_ptr);
int ba_host_selftest(uint32_t nb, uint32_t max_per_bucket, int rounds, uint32_t seed);
// a[i] <- 1 / a[i] over Fq (zeros stay zero); scratch >= n / 32 + 64 elements
void fq_batch_invert(Fq* a, size_t n, cudaStream_t st, LaunchCounter lc, Fq* scratch);

}  // namespace zg
