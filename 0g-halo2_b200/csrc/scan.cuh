// Multi-CTA exclusive scan of u32 counters (bucket offsets of the MSM counting sort, radix-sort digit
// offsets and the placement bookkeeping of the lookup permutation).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "poly.cuh"

namespace zg {

constexpr uint32_t SCAN_PER_THREAD = 8;
constexpr uint32_t SCAN_TILE = 1024 * SCAN_PER_THREAD;   // counters per CTA (a whole radix pass of the k = 15 lookups is one tile)
constexpr uint32_t SCAN_MAX_JOBS = 4;

struct ScanJobs {
  const uint32_t* in[SCAN_MAX_JOBS];
  uint32_t* out[SCAN_MAX_JOBS];    // n + 1 entries: out[n] = total
  uint32_t* out2[SCAN_MAX_JOBS];   // optional second copy of the n + 1 results (may be null)
};

// words of scratch needed for `njobs` scans of n counters
inline size_t scan_scratch_words(uint32_t n, uint32_t njobs) { return (size_t)njobs * ((n + SCAN_TILE - 1) / SCAN_TILE + 1); }

// njobs (<= 4) independent exclusive scans of n counters each.  One launch when n <= SCAN_TILE, else two.
void scan_excl_u32(const ScanJobs& jobs, uint32_t njobs, uint32_t n, uint32_t* scratch, cudaStream_t st, LaunchCounter lc);

}  // namespace zg
