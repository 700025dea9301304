// reduce.cuh -- compensated accumulators and fixed-order block reductions.
#pragma once
#include <cuda_runtime.h>

namespace mg {

// Neumaier compensated sum: value() carries ~2x float64 precision, so a
// parallel combination in any fixed order lands within an ulp or two of the
// exactly rounded sum.
struct Comp {
  double s = 0.0, c = 0.0;
  __host__ __device__ __forceinline__ void add(double x) {
    const double t = s + x;
    c += (fabs(s) >= fabs(x)) ? ((s - t) + x) : ((x - t) + s);
    s = t;
  }
  __host__ __device__ __forceinline__ void merge(const Comp &o) { add(o.s); c += o.c; }
  __host__ __device__ __forceinline__ double value() const { return s + c; }
};

__device__ __forceinline__ Comp warp_reduce_comp(Comp a) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Comp o;
    o.s = __shfl_down_sync(0xffffffffu, a.s, off);
    o.c = __shfl_down_sync(0xffffffffu, a.c, off);
    a.merge(o);
  }
  return a;
}

// Result valid in thread 0.  All threads of the block must call.
template <int BLOCK>
__device__ __forceinline__ double block_reduce_comp(Comp a) {
  __shared__ double sh_s[BLOCK / 32], sh_c[BLOCK / 32];
  a = warp_reduce_comp(a);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh_s[w] = a.s; sh_c[w] = a.c; }
  __syncthreads();
  Comp t;
  if (threadIdx.x == 0) {
    for (int k = 0; k < BLOCK / 32; ++k) { Comp o; o.s = sh_s[k]; o.c = sh_c[k]; t.merge(o); }
  }
  __syncthreads();
  return t.value();
}

}  // namespace mg
