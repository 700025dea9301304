// reduce_sum.cuh -- sum_i f(i) over a device index range with compensated
// per-thread partials and a fixed-order combination (deterministic, within an
// ulp or two of the exactly rounded sum).  f is an extended __device__ lambda.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace mg {

constexpr int EB = 256;

inline unsigned egrid(mg_ctx *ctx, int64_t n, int block = EB) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = (int64_t)ctx->sm_count * 8;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

// partial[b] = sum over this block's elements of f(i)
template <class F>
static __global__ void __launch_bounds__(EB) reduce_kernel(int64_t n, F f, double *__restrict__ partial) {
  Comp acc;
  for (int64_t i = (int64_t)blockIdx.x * EB + threadIdx.x; i < n; i += (int64_t)gridDim.x * EB) acc.add(f(i));
  const double t = block_reduce_comp<EB>(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

template <class F>
inline int reduce_sum(mg_ctx *ctx, int64_t n, F f, double *out) {
  const unsigned g = egrid(ctx, n);
  DevBuf<double> partial;
  MG_CUDA(ctx, partial.alloc(g, ctx->stream));
  reduce_kernel<<<g, EB, 0, ctx->stream>>>(n, f, partial.get());
  MG_CHECK_LAUNCH(ctx);
  std::vector<double> h(g);
  MG_CUDA(ctx, cudaMemcpyAsync(h.data(), partial.get(), sizeof(double) * g, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  Comp acc;
  for (unsigned b = 0; b < g; ++b) acc.add(h[b]);  // fixed order
  *out = acc.value();
  return MG_OK;
}

}  // namespace mg
