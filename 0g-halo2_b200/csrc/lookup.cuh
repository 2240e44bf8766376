// Lookup permutation (permute_expression_pair) on the device.
#pragma once
#include <cuda_runtime.h>
#include "poly.cuh"

namespace zg {

size_t lookup_workspace_bytes(uint32_t n);

// a_mont / s_mont: compressed input / table columns (Montgomery), first `usable` rows are permuted into
// pa / ps (Montgomery).  status_dev[0] is OR-ed with 1 when the partial (top 64 bits) sort left the table
// unsorted (caller retries with full_sort), status_dev[1] when an input value is missing from the table.
int lookup_permute(const Fr* a_mont, const Fr* s_mont, uint32_t usable, Fr* pa, Fr* ps, uint8_t* ws, uint32_t* status_dev,
                   bool full_sort, cudaStream_t st, LaunchCounter lc);

}  // namespace zg
