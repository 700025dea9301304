// peaks.cu -- device microbenchmarks used for the roofline denominators that
// MEASURED_PEAKS.json does not carry: FP64 FMA throughput and streaming-store
// bandwidth (the MH ensemble kernel only writes).
#include "common.cuh"

namespace mg {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) store_peak_kernel(double *out, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    __stcs(out + i, v);
}

}  // namespace mg

using namespace mg;

// FP64 FMA throughput in TFLOP/s (2 flops per DFMA), best of `reps`.
extern "C" int mg_measure_fp64_tflops(mg_ctx *ctx, int reps, double *out_tflops) {
  if (!ctx || !out_tflops) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, iters = 1 << 15;
  DevBuf<double> buf;
  MG_CUDA(ctx, buf.alloc((size_t)blocks * 256, ctx->stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0, ctx->stream);
    dfma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(buf.get(), iters, 1.0);
    MG_CHECK_LAUNCH(ctx);
    cudaEventRecord(e1, ctx->stream);
    MG_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *out_tflops = best;
  return MG_OK;
}

// Streaming-store bandwidth in GB/s over `nbytes` (>> L2), best of `reps`.
extern "C" int mg_measure_store_gbs(mg_ctx *ctx, int64_t nbytes, int reps, double *out_gbs) {
  if (!ctx || !out_gbs || nbytes < 8) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf<double> buf;
  const int64_t n = nbytes / 8;
  MG_CUDA(ctx, buf.alloc((size_t)n, ctx->stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0, ctx->stream);
    store_peak_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(buf.get(), n, (double)r);
    MG_CHECK_LAUNCH(ctx);
    cudaEventRecord(e1, ctx->stream);
    MG_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    const double gbs = (double)n * 8.0 / (ms * 1e-3) / 1e9;
    if (r > 0 && gbs > best) best = gbs;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *out_gbs = best;
  return MG_OK;
}
