// Lookup permutation (permute_expression_pair) on the device.
#pragma once
#include <cuda_runtime.h>
#include "poly.cuh"

namespace zg {

// sorted unique table values (canonical form) with run starts; D = number of distinct values (device)
struct LookupTable {
  Fr* U = nullptr;
  uint32_t* ustart = nullptr;
  uint32_t* D = nullptr;
};
size_t lookup_table_bytes(uint32_t n);
LookupTable lookup_table_carve(uint8_t* mem, uint32_t n);

size_t lookup_workspace_bytes(uint32_t n);
// scratch LookupTable inside a workspace of lookup_workspace_bytes(n)
LookupTable lookup_workspace_table(uint8_t* ws, uint32_t n);

// sort the first `usable` rows of the compressed table column (Montgomery) into `out`.
// *unsorted_flag_dev is OR-ed with 1 when the partial (top 64 bits) sort was not enough (retry full_sort).
int lookup_sort_table(const Fr* s_mont, uint32_t usable, const LookupTable& out, uint8_t* ws, uint32_t* unsorted_flag_dev,
                      bool full_sort, cudaStream_t st, LaunchCounter lc);
// permuted input / table columns (Montgomery) for the first `usable` rows.
// *missing_flag_dev is OR-ed with 1 when an input value does not occur in the table.
int lookup_permute(const Fr* a_mont, uint32_t usable, const LookupTable& tab, Fr* pa, Fr* ps, uint8_t* ws,
                   uint32_t* missing_flag_dev, cudaStream_t st, LaunchCounter lc);

}  // namespace zg
