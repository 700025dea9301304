// radix_sort.cuh -- stable LSD radix sort of (uint64 key, int32 value) pairs,
// batched over nb independent arrays of length n (blockIdx.y = array).
// 8 passes of 8-bit digits; every pass is
//   histogram (per 4096-key tile, smem atomics)
//   -> exclusive scan of the [digit][tile] table (scan.cuh)
//   -> stable scatter: each warp ranks its contiguous 512-key chunk with
//      match.any, warps are ordered inside the tile by a per-digit prefix.
// A pass in which every key has the same digit is skipped by the host.
// HBM-bound: 8 B read (histogram) + 12 B read + 12 B written (scatter) per
// pair per executed pass.
//
// Used by the kd-tree build (one sort per coordinate, kd_tree.ml:69-86's
// order statistics become array lookups) and by Evidence (sort by -ll,
// evidence.ml:180).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace mg {

constexpr int RS_BLOCK = 256;
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_ROUNDS = 16;                       // 32 keys per warp per round
constexpr int RS_TILE = RS_BLOCK * RS_ROUNDS;       // 4096 keys per CTA
constexpr int RS_RADIX = 256;

// float64 -> uint64 whose unsigned order is Pervasives.compare's order on
// non-NaN floats: -0.0 = +0.0 (kd_tree.ml:88-91 compares with `compare`).
__host__ __device__ __forceinline__ uint64_t f64_to_ordered(double x) {
  if (x == 0.0) x = 0.0;  // canonicalise -0.0
#ifdef __CUDA_ARCH__
  uint64_t b = (uint64_t)__double_as_longlong(x);
#else
  uint64_t b; memcpy(&b, &x, 8);
#endif
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

static __global__ void __launch_bounds__(RS_BLOCK)
rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int64_t ntiles, int shift,
               int32_t *__restrict__ hist /* [nb][256][ntiles] */) {
  __shared__ int sh[RS_RADIX];
  const int64_t b = blockIdx.y, t = blockIdx.x;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t *src = keys + b * n;
  const int64_t base = t * RS_TILE;
#pragma unroll 4
  for (int k = 0; k < RS_ROUNDS; ++k) {
    const int64_t i = base + k * RS_BLOCK + threadIdx.x;
    if (i < n) atomicAdd(&sh[(src[i] >> shift) & 0xFF], 1);
  }
  __syncthreads();
  hist[(b * RS_RADIX + threadIdx.x) * ntiles + t] = sh[threadIdx.x];
}

static __global__ void __launch_bounds__(RS_BLOCK)
rs_scatter_kernel(const uint64_t *__restrict__ keys_in, const int32_t *__restrict__ vals_in, int64_t n,
                  int64_t ntiles, int shift, const int32_t *__restrict__ offs /* scanned hist */,
                  uint64_t *__restrict__ keys_out, int32_t *__restrict__ vals_out) {
  __shared__ int cnt[RS_WARPS][RS_RADIX];   // per-warp digit counts, then per-warp bases
  const int64_t b = blockIdx.y, t = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < RS_WARPS * RS_RADIX; k += RS_BLOCK) (&cnt[0][0])[k] = 0;
  __syncthreads();
  const uint64_t *ksrc = keys_in + b * n;
  const int32_t *vsrc = vals_in + b * n;
  // warp w owns keys [base + w*512, base + (w+1)*512), 16 rounds of 32 consecutive keys
  const int64_t wbase = t * RS_TILE + (int64_t)w * (32 * RS_ROUNDS);
  uint64_t key[RS_ROUNDS];
  int rank[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
    key[r] = valid ? ksrc[i] : ~0ull;
    const int d = (int)((key[r] >> shift) & 0xFF);
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 256);
    const int before = __popc(peers & ((1u << lane) - 1u));
    int old = 0;
    if (valid) old = cnt[w][d];
    __syncwarp();
    if (valid && before == 0) cnt[w][d] = old + __popc(peers);
    __syncwarp();
    rank[r] = old + before;
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps of this tile + global base
  {
    const int d = threadIdx.x;
    int run = offs[(b * RS_RADIX + d) * ntiles + t];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) { const int c = cnt[ww][d]; cnt[ww][d] = run; run += c; }
  }
  __syncthreads();
  uint64_t *kdst = keys_out + b * n;
  int32_t *vdst = vals_out + b * n;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const int d = (int)((key[r] >> shift) & 0xFF);
      const int64_t pos = (int64_t)cnt[w][d] + rank[r];
      kdst[pos] = key[r];
      vdst[pos] = vsrc[i];
    }
  }
}

struct RadixSortTemp {
  DevBuf<uint64_t> keys_alt;
  DevBuf<int32_t> vals_alt, hist, scan_tmp, totals;
  std::vector<int32_t> h_hist;
};

// Sorts in place (result ends in d_keys / d_vals).  key_bits: number of
// significant low bits (64 for doubles).  Returns MG_OK or an error.
inline int radix_sort_pairs(mg_ctx *ctx, uint64_t *d_keys, int32_t *d_vals, int64_t n, int64_t nb, int key_bits = 64) {
  if (n <= 1 || nb <= 0) return MG_OK;
  cudaStream_t s = ctx->stream;
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  const int64_t hist_n = RS_RADIX * ntiles;  // per batch row
  if (hist_n >= 2147483647LL) return set_err(ctx, MG_EINVAL, "radix sort: array too long");
  RadixSortTemp tmp;
  MG_CUDA(ctx, tmp.keys_alt.alloc((size_t)n * nb, s));
  MG_CUDA(ctx, tmp.vals_alt.alloc((size_t)n * nb, s));
  MG_CUDA(ctx, tmp.hist.alloc((size_t)hist_n * nb, s));
  MG_CUDA(ctx, tmp.scan_tmp.alloc((size_t)scan_tmp_elems(hist_n, nb), s));
  uint64_t *kin = d_keys, *kout = tmp.keys_alt.get();
  int32_t *vin = d_vals, *vout = tmp.vals_alt.get();
  dim3 grid((unsigned)ntiles, (unsigned)nb);
  for (int shift = 0; shift < key_bits; shift += 8) {
    rs_hist_kernel<<<grid, RS_BLOCK, 0, s>>>(kin, n, ntiles, shift, tmp.hist.get());
    MG_CHECK_LAUNCH(ctx);
    int rc = exclusive_scan_i32(ctx, tmp.hist.get(), tmp.hist.get(), hist_n, nb, tmp.scan_tmp.get(), nullptr);
    if (rc) return rc;
    rs_scatter_kernel<<<grid, RS_BLOCK, 0, s>>>(kin, vin, n, ntiles, shift, tmp.hist.get(), kout, vout);
    MG_CHECK_LAUNCH(ctx);
    std::swap(kin, kout); std::swap(vin, vout);
  }
  if (kin != d_keys) {  // odd number of passes: copy back
    MG_CUDA(ctx, cudaMemcpyAsync(d_keys, kin, sizeof(uint64_t) * n * nb, cudaMemcpyDeviceToDevice, s));
    MG_CUDA(ctx, cudaMemcpyAsync(d_vals, vin, sizeof(int32_t) * n * nb, cudaMemcpyDeviceToDevice, s));
  }
  return MG_OK;
}

}  // namespace mg
