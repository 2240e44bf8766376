// Multi-CTA exclusive scan of u32 counters: tile-local scan + tile totals, then every tile adds the sum of
// the totals before it.  Replaces the single-CTA scans that sat on the critical path of every MSM batch and of
// every radix-sort pass (ncu launch lists profiles/r01_launches_proof_large_*.csv: 3.3 ms of a 30 ms proof).
#include "scan.cuh"

namespace zg {

namespace {

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane >= d) incl += o;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
      if ((int)lane >= d) wi += o;
    }
    warp_sums[lane] = wi - ws;
    if (lane == 31) warp_sums[32] = wi;
  }
  __syncthreads();
  total = warp_sums[32];
  return warp_sums[wid] + incl - v;
}

// grid (tiles, njobs).  Tile-local exclusive scan; with a single tile the result is final.
__global__ void __launch_bounds__(1024) k_scan_tiles(ScanJobs J, uint32_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t warp_sums[33];
  const uint32_t job = blockIdx.y, tiles = gridDim.x;
  const uint32_t* in = J.in[job];
  uint32_t* out = J.out[job];
  uint32_t* out2 = J.out2[job];
  constexpr int E = (int)SCAN_PER_THREAD;
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * E;
  uint32_t v[E];
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < E; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    sum += v[i];
  }
  uint32_t total;
  uint32_t run = block_excl_scan_1024(sum, warp_sums, total);
  const bool final_pass = tiles == 1;
#pragma unroll
  for (int i = 0; i < E; i++) {
    if (base + i < n) {
      out[base + i] = run;
      if (final_pass && out2) out2[base + i] = run;
    }
    run += v[i];
  }
  if (threadIdx.x == 0) {
    if (final_pass) {
      out[n] = total;
      if (out2) out2[n] = total;
    } else {
      tile_sums[job * tiles + blockIdx.x] = total;
    }
  }
}

// grid (tiles, njobs): add the totals of the tiles before this one
__global__ void __launch_bounds__(1024) k_scan_add_base(ScanJobs J, uint32_t n, const uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t warp_sums[33];
  const uint32_t job = blockIdx.y, tiles = gridDim.x, tile = blockIdx.x;
  uint32_t* out = J.out[job];
  uint32_t* out2 = J.out2[job];
  const uint32_t* ts = tile_sums + job * tiles;
  uint32_t part = 0;
  for (uint32_t b = threadIdx.x; b < tile; b += 1024) part += ts[b];
  uint32_t total;
  block_excl_scan_1024(part, warp_sums, total);   // total = sum of the tiles before this one
  const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
#pragma unroll
  for (int i = 0; i < (int)SCAN_PER_THREAD; i++) {
    if (base + i < n) {
      uint32_t r = out[base + i] + total;
      out[base + i] = r;
      if (out2) out2[base + i] = r;
    }
  }
  if (tile == tiles - 1 && threadIdx.x == 0) {
    uint32_t all = total + ts[tile];
    out[n] = all;
    if (out2) out2[n] = all;
  }
}

}  // namespace

void scan_excl_u32(const ScanJobs& jobs, uint32_t njobs, uint32_t n, uint32_t* scratch, cudaStream_t st, LaunchCounter lc) {
  if (njobs == 0) return;
  const uint32_t tiles = n == 0 ? 1 : (n + SCAN_TILE - 1) / SCAN_TILE;
  dim3 grid(tiles, njobs);
  k_scan_tiles<<<grid, 1024, 0, st>>>(jobs, n, scratch);
  lc++;
  if (tiles > 1) {
    k_scan_add_base<<<grid, 1024, 0, st>>>(jobs, n, scratch);
    lc++;
  }
}

}  // namespace zg
