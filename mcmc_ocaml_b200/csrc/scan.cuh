// scan.cuh -- batched exclusive prefix sums over int32 arrays [nb][n]
// (reduce tiles -> scan tile sums -> downsweep), the building block of the
// radix sort and of the kd-tree's stable partitions.  HBM-bound: 4 B read in
// the reduce pass, 4 B read + 4 B written in the downsweep.
#pragma once
#include "common.cuh"

namespace mg {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;  // 2048

// block-wide exclusive scan of one int per thread; returns exclusive prefix,
// total in *total (valid in all threads)
template <int BLOCK>
__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[BLOCK / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = (lane < BLOCK / 32) ? warp_sums[lane] : 0;
    int si = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, si, off);
      if (lane >= off) si += t;
    }
    if (lane < BLOCK / 32) warp_sums[lane] = si - s;
    if (lane == BLOCK / 32 - 1) block_total = si;
  }
  __syncthreads();
  const int res = incl - v + warp_sums[w];
  *total = block_total;
  __syncthreads();
  return res;
}

// tile sums: sums[b][t] = sum in[b][t*TILE .. (t+1)*TILE)
static __global__ void __launch_bounds__(SCAN_BLOCK)
scan_reduce_kernel(const int32_t *__restrict__ in, int64_t n, int64_t ntiles, int32_t *__restrict__ sums) {
  const int64_t b = blockIdx.y, t = blockIdx.x;
  const int32_t *src = in + b * n;
  int acc = 0;
  const int64_t base = t * SCAN_TILE;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k * SCAN_BLOCK + threadIdx.x;
    if (i < n) acc += src[i];
  }
  int total;
  block_exclusive_scan<SCAN_BLOCK>(acc, &total);
  if (threadIdx.x == 0) sums[b * ntiles + t] = total;
}

// in-place exclusive scan of sums[b][0..ntiles) by one block per batch row;
// also writes the row total to totals[b] if non-null
static __global__ void __launch_bounds__(1024)
scan_tiles_kernel(int32_t *__restrict__ sums, int64_t ntiles, int32_t *__restrict__ totals) {
  const int64_t b = blockIdx.x;
  int32_t *row = sums + b * ntiles;
  __shared__ int carry_sh;
  if (threadIdx.x == 0) carry_sh = 0;
  __syncthreads();
  for (int64_t base = 0; base < ntiles; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = (i < ntiles) ? row[i] : 0;
    int total;
    const int ex = block_exclusive_scan<1024>(v, &total);
    const int carry = carry_sh;
    if (i < ntiles) row[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_sh = carry + total;
    __syncthreads();
  }
  if (totals && threadIdx.x == 0) totals[b] = carry_sh;
}

// out[b][i] = tile_base[b][t] + exclusive prefix within the tile
static __global__ void __launch_bounds__(SCAN_BLOCK)
scan_down_kernel(const int32_t *__restrict__ in, int64_t n, int64_t ntiles, const int32_t *__restrict__ sums,
                 int32_t *__restrict__ out) {
  const int64_t b = blockIdx.y, t = blockIdx.x;
  const int32_t *src = in + b * n;
  int32_t *dst = out + b * n;
  // blocked arrangement: thread owns ITEMS consecutive elements
  const int64_t base = t * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int acc = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? src[base + k] : 0; acc += v[k]; }
  int total;
  int ex = block_exclusive_scan<SCAN_BLOCK>(acc, &total) + sums[b * ntiles + t];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) dst[base + k] = ex; ex += v[k]; }
}

// Exclusive scan of nb rows of n int32 each.  d_tmp must hold nb * ceil(n/TILE) int32.
inline int64_t scan_tmp_elems(int64_t n, int64_t nb) { return nb * ((n + SCAN_TILE - 1) / SCAN_TILE); }

static inline int exclusive_scan_i32(mg_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t n, int64_t nb,
                              int32_t *d_tmp, int32_t *d_totals) {
  if (n <= 0 || nb <= 0) return MG_OK;
  const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  dim3 grid((unsigned)ntiles, (unsigned)nb);
  scan_reduce_kernel<<<grid, SCAN_BLOCK, 0, ctx->stream>>>(d_in, n, ntiles, d_tmp);
  MG_CHECK_LAUNCH(ctx);
  scan_tiles_kernel<<<(unsigned)nb, 1024, 0, ctx->stream>>>(d_tmp, ntiles, d_totals);
  MG_CHECK_LAUNCH(ctx);
  scan_down_kernel<<<grid, SCAN_BLOCK, 0, ctx->stream>>>(d_in, n, ntiles, d_tmp, d_out);
  MG_CHECK_LAUNCH(ctx);
  return MG_OK;
}

}  // namespace mg
