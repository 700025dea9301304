// radix_sort.cuh -- stable LSD radix sort of (uint64 key, int32 value) pairs,
// batched over nb independent arrays of length n (blockIdx.y = array).
// 8 passes of 8-bit digits; every pass is
//   histogram (per 4096-key tile, smem atomics)
//   -> exclusive scan of the [digit][tile] table (scan.cuh)
//   -> stable scatter: each warp ranks its contiguous 512-key chunk with
//      match.any, warps are ordered inside the tile by a per-digit prefix.
// A pass in which every key has the same digit is skipped (one up-front pass
// builds all eight global digit histograms).
// HBM-bound: 8 B read (histogram) + 12 B read + 12 B written (scatter) per
// pair per executed pass.
//
// Used by the kd-tree build (one sort per coordinate, kd_tree.ml:69-86's
// order statistics become array lookups) and by Evidence (sort by -ll,
// evidence.ml:180).
#pragma once
#include "common.cuh"
#include "scan.cuh"

#include <vector>

namespace mg {

#ifndef MG_RS_MINBLOCKS
#define MG_RS_MINBLOCKS 4
#endif
constexpr int RS_BLOCK = 256;
constexpr int RS_WARPS = RS_BLOCK / 32;
#ifndef MG_RS_ROUNDS
#define MG_RS_ROUNDS 8
#endif
constexpr int RS_ROUNDS = MG_RS_ROUNDS;             // 32 keys per warp per round
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

template <class KeyT>
static __global__ void __launch_bounds__(RS_BLOCK)
rs_hist_kernel(const KeyT *__restrict__ keys, int64_t n, int64_t ntiles, int shift,
               int32_t *__restrict__ hist /* [nb][256][ntiles] */) {
  __shared__ int sh[RS_RADIX];
  const int64_t b = blockIdx.y, t = blockIdx.x;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const KeyT *src = keys + b * n;
  const int64_t base = t * RS_TILE;
#pragma unroll 4
  for (int k = 0; k < RS_ROUNDS; ++k) {
    const int64_t i = base + k * RS_BLOCK + threadIdx.x;
    if (i < n) atomicAdd(&sh[(src[i] >> shift) & 0xFF], 1);
  }
  __syncthreads();
  hist[(b * RS_RADIX + threadIdx.x) * ntiles + t] = sh[threadIdx.x];
}

// Stable scatter of one 4096-key tile.  Ranks: each warp walks its contiguous
// 512-key chunk in 16 rounds of 32 consecutive keys and ranks equal digits
// with match.any (stable by construction); warps are ordered by a per-digit
// prefix.  The tile is then written to shared memory IN DIGIT ORDER and copied
// out from there, so that every digit's run leaves as contiguous, coalesced
// global stores (a direct scatter writes 8-byte keys to 256 different streams:
// one 32-byte sector per key).
template <class KeyT>
constexpr size_t rs_scatter_smem() {
  return (size_t)RS_TILE * (sizeof(KeyT) + sizeof(int32_t)) + (size_t)(RS_WARPS * RS_RADIX + 2 * RS_RADIX) * sizeof(int);
}

template <class KeyT>
static __global__ void __launch_bounds__(RS_BLOCK, MG_RS_MINBLOCKS)
rs_scatter_kernel(const KeyT *__restrict__ keys_in, const int32_t *__restrict__ vals_in, int64_t n,
                  int64_t ntiles, int shift, const int32_t *__restrict__ offs /* scanned hist */,
                  KeyT *__restrict__ keys_out, int32_t *__restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  KeyT *skeys = reinterpret_cast<KeyT *>(rs_smem);
  int32_t *svals = reinterpret_cast<int32_t *>(skeys + RS_TILE);
  int (*cnt)[RS_RADIX] = reinterpret_cast<int (*)[RS_RADIX]>(svals + RS_TILE);  // [RS_WARPS][RS_RADIX]
  int *dbase = &cnt[RS_WARPS][0];   // digit start inside the tile
  int *gdelta = dbase + RS_RADIX;   // global offset of the digit minus dbase
  const int64_t b = blockIdx.y, t = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < RS_WARPS * RS_RADIX; k += RS_BLOCK) (&cnt[0][0])[k] = 0;
  __syncthreads();
  const KeyT *ksrc = keys_in + b * n;
  const int32_t *vsrc = vals_in + b * n;
  const int64_t tile0 = t * RS_TILE;
  const int64_t wbase = tile0 + (int64_t)w * (32 * RS_ROUNDS);
  const int nvalid = (int)((n - tile0 < RS_TILE) ? (n - tile0) : RS_TILE);
  KeyT key[RS_ROUNDS];
  int32_t val[RS_ROUNDS];
  int rank[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {   // all the tile's loads in flight before the first ranking round
    const int64_t i = wbase + r * 32 + lane;
    key[r] = (i < n) ? __ldcs(ksrc + i) : (KeyT)~(KeyT)0;
    val[r] = (i < n) ? __ldcs(vsrc + i) : 0;
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
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
  {
    // thread d: prefix over warps, then block-wide exclusive scan over digits
    const int d = threadIdx.x;
    int run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) { const int c = cnt[ww][d]; cnt[ww][d] = run; run += c; }
    int total;
    const int start = block_exclusive_scan<RS_BLOCK>(run, &total);
    dbase[d] = start;
    gdelta[d] = offs[(b * RS_RADIX + d) * ntiles + t] - start;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const int d = (int)((key[r] >> shift) & 0xFF);
      const int pos = dbase[d] + cnt[w][d] + rank[r];
      skeys[pos] = key[r];
      svals[pos] = val[r];
    }
  }
  __syncthreads();
  KeyT *kdst = keys_out + b * n;
  int32_t *vdst = vals_out + b * n;
  for (int i = threadIdx.x; i < nvalid; i += RS_BLOCK) {
    const KeyT k = skeys[i];
    const int64_t pos = (int64_t)gdelta[(int)((k >> shift) & 0xFF)] + i;
    kdst[pos] = k;
    vdst[pos] = svals[i];
  }
}

// all global digit histograms (one per key byte) in one pass over the keys: [nb][sizeof(KeyT)][256]
template <class KeyT>
static __global__ void __launch_bounds__(RS_BLOCK)
rs_prehist_kernel(const KeyT *__restrict__ keys, int64_t n, unsigned int *__restrict__ ghist) {
  constexpr int NB = (int)sizeof(KeyT);
  __shared__ unsigned int sh[NB][RS_RADIX];
  const int64_t b = blockIdx.y;
  for (int k = threadIdx.x; k < NB * RS_RADIX; k += RS_BLOCK) (&sh[0][0])[k] = 0u;
  __syncthreads();
  const KeyT *src = keys + b * n;
  for (int64_t i = (int64_t)blockIdx.x * RS_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_BLOCK) {
    const KeyT k = src[i];
#pragma unroll
    for (int p = 0; p < NB; ++p) atomicAdd(&sh[p][(k >> (8 * p)) & 0xFF], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < NB * RS_RADIX; k += RS_BLOCK) {
    const unsigned int v = (&sh[0][0])[k];
    if (v) atomicAdd(ghist + b * NB * RS_RADIX + k, v);
  }
}

// const_byte[b][p] = 1 iff byte p of every key of row b is the same (a radix pass over it moves nothing)
template <class KeyT>
static inline int radix_constant_bytes(mg_ctx *ctx, const KeyT *d_keys, int64_t n, int64_t nb, std::vector<unsigned char> &const_byte,
                                       std::vector<double> *eff_values = nullptr /* [nb][bytes]: n^2 / sum count^2 */) {
  constexpr int NB = (int)sizeof(KeyT);
  cudaStream_t s = ctx->stream;
  DevBuf<unsigned int> ghist;
  MG_CUDA(ctx, ghist.alloc((size_t)nb * NB * RS_RADIX, s));
  MG_CUDA(ctx, cudaMemsetAsync(ghist.get(), 0, sizeof(unsigned int) * nb * NB * RS_RADIX, s));
  int64_t gx = (n + RS_BLOCK * 8 - 1) / (RS_BLOCK * 8);
  if (gx > (int64_t)ctx->sm_count * 8) gx = (int64_t)ctx->sm_count * 8;
  rs_prehist_kernel<KeyT><<<dim3((unsigned)gx, (unsigned)nb), RS_BLOCK, 0, s>>>(d_keys, n, ghist.get());
  MG_CHECK_LAUNCH(ctx);
  std::vector<unsigned int> h((size_t)nb * NB * RS_RADIX);
  MG_CUDA(ctx, cudaMemcpyAsync(h.data(), ghist.get(), sizeof(unsigned int) * h.size(), cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  const_byte.assign((size_t)nb * NB, 0);
  if (eff_values) eff_values->assign((size_t)nb * NB, 1.0);
  for (int64_t b = 0; b < nb; ++b)
    for (int p = 0; p < NB; ++p) {
      double sq = 0.0;
      for (int d = 0; d < RS_RADIX; ++d) {
        const unsigned int c = h[((size_t)b * NB + p) * RS_RADIX + d];
        if (c == (unsigned int)n) const_byte[(size_t)b * NB + p] = 1;
        sq += (double)c * (double)c;
      }
      if (eff_values && sq > 0.0) (*eff_values)[(size_t)b * NB + p] = (double)n * (double)n / sq;
    }
  return MG_OK;
}

template <class KeyT>
struct RadixSortTemp {
  DevBuf<KeyT> keys_alt;
  DevBuf<int32_t> vals_alt, hist, scan_tmp, totals;
  std::vector<int32_t> h_hist;
};

// Sorts in place (result ends in d_keys / d_vals).  key_bits: number of
// significant low bits (the key width by default).  Returns MG_OK or an error.
template <class KeyT>
static inline int radix_sort_pairs_t(mg_ctx *ctx, KeyT *d_keys, int32_t *d_vals, int64_t n, int64_t nb,
                                     int key_bits = 8 * (int)sizeof(KeyT)) {
  constexpr int NB = (int)sizeof(KeyT);
  if (n <= 1 || nb <= 0) return MG_OK;
  cudaStream_t s = ctx->stream;
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  const int64_t hist_n = RS_RADIX * ntiles;  // per batch row
  if (hist_n >= 2147483647LL) return set_err(ctx, MG_EINVAL, "radix sort: array too long");
  RadixSortTemp<KeyT> tmp;
  MG_CUDA(ctx, tmp.keys_alt.alloc((size_t)n * nb, s));
  MG_CUDA(ctx, tmp.vals_alt.alloc((size_t)n * nb, s));
  MG_CUDA(ctx, tmp.hist.alloc((size_t)hist_n * nb, s));
  MG_CUDA(ctx, tmp.scan_tmp.alloc((size_t)scan_tmp_elems(hist_n, nb), s));
  KeyT *kin = d_keys, *kout = tmp.keys_alt.get();
  int32_t *vin = d_vals, *vout = tmp.vals_alt.get();
  dim3 grid((unsigned)ntiles, (unsigned)nb);
  // a pass whose digit is the same for every key of every row moves nothing: find those up front
  bool skip[NB];
  for (int p = 0; p < NB; ++p) skip[p] = false;
  if (n >= 4 * RS_TILE) {
    std::vector<unsigned char> cb;
    int rc = radix_constant_bytes<KeyT>(ctx, d_keys, n, nb, cb);
    if (rc) return rc;
    for (int p = 0; p < NB; ++p) {
      bool all_const = true;
      for (int64_t b = 0; b < nb && all_const; ++b) all_const = cb[(size_t)b * NB + p] != 0;
      skip[p] = all_const;
    }
  }
  MG_CUDA(ctx, cudaFuncSetAttribute(rs_scatter_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_scatter_smem<KeyT>()));
  for (int shift = 0; shift < key_bits; shift += 8) {
    if (skip[shift / 8]) continue;
    rs_hist_kernel<KeyT><<<grid, RS_BLOCK, 0, s>>>(kin, n, ntiles, shift, tmp.hist.get());
    MG_CHECK_LAUNCH(ctx);
    int rc = exclusive_scan_i32(ctx, tmp.hist.get(), tmp.hist.get(), hist_n, nb, tmp.scan_tmp.get(), nullptr);
    if (rc) return rc;
    rs_scatter_kernel<KeyT><<<grid, RS_BLOCK, rs_scatter_smem<KeyT>(), s>>>(kin, vin, n, ntiles, shift, tmp.hist.get(), kout, vout);
    MG_CHECK_LAUNCH(ctx);
    std::swap(kin, kout); std::swap(vin, vout);
  }
  if (kin != d_keys) {  // odd number of passes: copy back
    MG_CUDA(ctx, cudaMemcpyAsync(d_keys, kin, sizeof(KeyT) * n * nb, cudaMemcpyDeviceToDevice, s));
    MG_CUDA(ctx, cudaMemcpyAsync(d_vals, vin, sizeof(int32_t) * n * nb, cudaMemcpyDeviceToDevice, s));
  }
  return MG_OK;
}

static inline int radix_sort_pairs(mg_ctx *ctx, uint64_t *d_keys, int32_t *d_vals, int64_t n, int64_t nb, int key_bits = 64) {
  return radix_sort_pairs_t<uint64_t>(ctx, d_keys, d_vals, n, nb, key_bits);
}

}  // namespace mg
