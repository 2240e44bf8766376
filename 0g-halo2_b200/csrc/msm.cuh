// MSM descriptors shared by msm.cu and the context layer.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "curve.cuh"

namespace zg {

// Fixed-base window table: pts[w * n + i] = 2^(c*w) * base[i], affine.
struct MsmTable {
  G1Affine* pts = nullptr;
  uint32_t n = 0, c = 0, W = 0;
};

constexpr uint32_t MSM_INVALID_KEY = 0xffffffffu;
// entries per thread of the second serial level: a chain of this many dependent additions per batch, so kept short
constexpr uint32_t MSM_LEVEL1_K = 8;

inline uint32_t msm_windows(uint32_t c) { return (255 + c - 1) / c; }
// window width used for a 2^k-point fixed-base MSM (overridable with ZG_MSM_C)
uint32_t msm_pick_c(uint32_t k);

struct MsmWorkspaceLayout {
  size_t bytes = 0;
  size_t off_hist, off_cursor, off_offsets, off_keys, off_vals, off_buckets;
  size_t off_pkeys_a, off_ppts_a, off_pkeys_b, off_ppts_b, off_s1, off_t1, off_l2, off_scan;
  uint32_t L_max, K0, T0, slots_a, slots_b, NB, M;
  uint32_t J, chunk;   // digit sort: J chunks of `chunk` scalars per MSM
};
MsmWorkspaceLayout msm_workspace_layout(uint32_t n, uint32_t c, uint32_t W, uint32_t M);

// Optional live timing of the level-0 accumulation kernel (the dominant kernel of the path): every launch is bracketed
// by CUDA events on the launching stream and its entry count (= mixed point additions) is copied back.
struct MsmProbe {
  bool on = false;
  std::vector<cudaEvent_t> ev;      // 2 per launch
  uint32_t* counts = nullptr;       // pinned host, one per launch
  size_t cap = 0, used = 0;
};

cudaError_t msm_precompute_table(const G1Affine* base, uint32_t n, uint32_t c, uint32_t W,
                                 G1Affine* table, cudaStream_t stream);

// M MSMs over the same table; scalars are Montgomery-form Fr on the device,
// polynomial m starts at scalars + m * scalar_stride.  `n_used` <= table.n scalars per MSM.
// out[m] is Jacobian.  `ws` must hold msm_workspace_layout(...).bytes.
cudaError_t msm_run(const MsmTable& table, const Fr* scalars, size_t scalar_stride, uint32_t n_used,
                    uint32_t M, G1Jac* out, uint8_t* ws, const MsmWorkspaceLayout& lay,
                    cudaStream_t stream, uint64_t* launch_counter, MsmProbe* probe = nullptr,
                    const G1Affine* alt_pts = nullptr, uint32_t alt_mask = 0);
// (alt_pts / alt_mask: MSM m of the batch reads the window table `alt_pts` instead when bit m of alt_mask is set --
//  one Fiat-Shamir round can mix commit_lagrange and commit columns in a single pipeline; both tables share n, c, W)

}  // namespace zg
