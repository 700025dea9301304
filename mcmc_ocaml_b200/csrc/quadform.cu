// quadform.cu -- the batched quadratic form  logc - 1/2 |L (x - mu)|^2  of a high-dimensional correlated Gaussian
// (MG_FN_GAUSS_CORR, the 32- / 64-D models of BASELINE.json config 5; generalises Stats.log_multi_gaussian,
// stats.ml:103-108) evaluated two ways, to decide with measurements whether the FP64 tensor cores are worth a
// likelihood variant of their own (north star: "tensor cores only for the batched quadratic forms"):
//   variant 0  one thread per point, D (D + 1) / 2 fused multiply-adds in the fixed order of the sampler's plugin
//              (bit-identical to MG_FN_GAUSS_CORR through mg_logfn_eval);
//   variant 1  one warp per 32 points: Y = L Z as an (D x D) x (D x 32) product on the FP64 tensor cores,
//              mma.sync.aligned.m8n8k4.f64 (DMMA; tcgen05 has no FP64 kind), zero tiles above the diagonal skipped,
//              Z and L staged in shared memory, squares and the reduction over rows on the fragments.
// The DMMA accumulates each 4-term dot product in an unspecified internal order, so variant 1 agrees with variant 0
// to a few ulp (stated tolerance 1e-13 relative in tests/test_misc_gpu.py), not bit for bit.
#include "common.cuh"
#include "models.cuh"

namespace mg {

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// peak: independent accumulator chains of DMMAs
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double seed) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) { c[k][0] = seed + k; c[k][1] = seed - k; }
  const double a = 0.999999 + 1e-9 * threadIdx.x, b = 1.000001;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) dmma_m8n8k4(c[k][0], c[k][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// variant 0: thread per point, parameters from shared memory (dense lower triangle, row major)
template <int D>
__global__ void __launch_bounds__(128) quadform_fma_kernel(const double *__restrict__ mu, const double *__restrict__ Lp, double logc,
                                                           const double *__restrict__ x, int64_t M, double *__restrict__ out) {
  __shared__ double sL[D * (D + 1) / 2], smu[D];
  for (int k = threadIdx.x; k < D * (D + 1) / 2; k += blockDim.x) sL[k] = Lp[k];
  for (int k = threadIdx.x; k < D; k += blockDim.x) smu[k] = mu[k];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    // parameters through volatile shared-memory loads, as the sampler's static plugin reads them (models.cuh lds1):
    // hoisting D (D + 1) / 2 loop-invariant values out of the point loop would spill them all
    double z[D];
#pragma unroll
    for (int j = 0; j < D; ++j) z[j] = x[(int64_t)j * M + i] - lds1(smu + j);    // [D][M]: coalesced, the sampler's state layout
    double q = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double y = lds1(sL + r * (r + 1) / 2) * z[0];
#pragma unroll
      for (int j = 1; j <= r; ++j) y = fma(lds1(sL + r * (r + 1) / 2 + j), z[j], y);
      q = fma(y, y, q);
    }
    out[i] = fma(-0.5, q, logc);
  }
}

// variant 1: warp per 32 points on the FP64 tensor cores
template <int D>
__global__ void __launch_bounds__(128) quadform_dmma_kernel(const double *__restrict__ mu, const double *__restrict__ Lp, double logc,
                                                            const double *__restrict__ x, int64_t M, double *__restrict__ out) {
  extern __shared__ double qf_smem[];
  constexpr int ZP = 33;                       // padded row of Z: [D][33]
  double *sL = qf_smem;                        // dense [D][D], zeros above the diagonal
  double *smu = sL + D * D;
  double *sZ = smu + D + (threadIdx.x >> 5) * (D * ZP);
  for (int k = threadIdx.x; k < D * D; k += blockDim.x) { const int r = k / D, c = k % D; sL[k] = c <= r ? Lp[r * (r + 1) / 2 + c] : 0.0; }
  for (int k = threadIdx.x; k < D; k += blockDim.x) smu[k] = mu[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int ar = lane >> 2, ac = lane & 3;     // fragment coordinates: A[ar][ac], B[k = ac][n = ar], C[ar][2 ac + {0,1}]
  for (int64_t w = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w * 32 < M; w += nwarps) {
    const int64_t i = w * 32 + lane;
    __syncwarp();
    for (int j = 0; j < D; ++j) sZ[j * ZP + lane] = (i < M) ? x[(int64_t)j * M + i] - smu[j] : 0.0;   // lane = point, [D][M] input
    __syncwarp();
    double q[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
    for (int mt = 0; mt < D / 8; ++mt) {
      double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
      for (int kk = 0; kk <= (8 * mt + 7) / 4; ++kk) {   // tiles at or below the diagonal
        const double a = sL[(8 * mt + ar) * D + 4 * kk + ac];
#pragma unroll
        for (int t = 0; t < 4; ++t) dmma_m8n8k4(acc[t][0], acc[t][1], a, sZ[(4 * kk + ac) * ZP + 8 * t + ar]);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) { q[t][0] = fma(acc[t][0], acc[t][0], q[t][0]); q[t][1] = fma(acc[t][1], acc[t][1], q[t][1]); }
    }
    // rows live on the lanes with equal (lane & 3): sum over lane bits 2..4
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v = q[t][h];
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        q[t][h] = v;
      }
    if (ar == 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t p = w * 32 + 8 * t + 2 * ac + h;
          if (p < M) out[p] = fma(-0.5, q[t][h], logc);
        }
    }
  }
}

}  // namespace mg

using namespace mg;

// FP64 tensor-core (DMMA m8n8k4) throughput in TFLOP/s (512 flops per warp instruction), best of `reps`.
extern "C" int mg_measure_dmma_tflops(mg_ctx *ctx, int reps, double *out_tflops) {
  if (!ctx || !out_tflops) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, iters = 1 << 13;
  DevBuf<double> buf;
  MG_CUDA(ctx, buf.alloc((size_t)blocks * 256, ctx->stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0, ctx->stream);
    dmma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(buf.get(), iters, 1.0);
    MG_CHECK_LAUNCH(ctx);
    cudaEventRecord(e1, ctx->stream);
    MG_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 512.0 * 8.0 * iters * 8.0 /* warps per block */ * blocks / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *out_tflops = best;
  return MG_OK;
}

// Diagnostic / decision experiment: the quadratic form of M device points [D][M] (the sampler's state layout) by
// variant 0 (FMA) or 1 (DMMA);
// D in {32, 64}.  *ms receives the kernel time (events on the context's stream, best of `reps`).
extern "C" int mg_debug_quadform(mg_ctx *ctx, int32_t variant, int32_t D, const double *mu, const double *Lpacked, double logc,
                                 const double *d_x, int64_t M, double *d_out, int32_t reps, double *ms) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, (D == 32 || D == 64) && (variant == 0 || variant == 1) && mu && Lpacked && d_x && d_out && M >= 1, "quadform: bad arguments");
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf<double> d_mu, d_L;
  MG_CUDA(ctx, upload(d_mu, mu, (size_t)D, s));
  MG_CUDA(ctx, upload(d_L, Lpacked, (size_t)D * (D + 1) / 2, s));
  const int block = 128;
  const unsigned grid = (unsigned)std::min<int64_t>((variant == 0 ? (M + block - 1) / block : (M + 31) / 32 / 4 + 1), (int64_t)ctx->sm_count * 16);
  const size_t smem = (size_t)(D * D + D + 4 * D * 33) * sizeof(double);
  if (variant == 1) {
    if (D == 32) MG_CUDA(ctx, cudaFuncSetAttribute(quadform_dmma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else MG_CUDA(ctx, cudaFuncSetAttribute(quadform_dmma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < std::max(1, reps) + 1; ++r) {
    cudaEventRecord(e0, s);
    if (variant == 0) {
      if (D == 32) quadform_fma_kernel<32><<<grid, block, 0, s>>>(d_mu.get(), d_L.get(), logc, d_x, M, d_out);
      else quadform_fma_kernel<64><<<grid, block, 0, s>>>(d_mu.get(), d_L.get(), logc, d_x, M, d_out);
    } else {
      if (D == 32) quadform_dmma_kernel<32><<<grid, block, smem, s>>>(d_mu.get(), d_L.get(), logc, d_x, M, d_out);
      else quadform_dmma_kernel<64><<<grid, block, smem, s>>>(d_mu.get(), d_L.get(), logc, d_x, M, d_out);
    }
    MG_CHECK_LAUNCH(ctx);
    cudaEventRecord(e1, s);
    MG_CUDA(ctx, cudaEventSynchronize(e1));
    float t = 0.f; cudaEventElapsedTime(&t, e0, e1);
    if (r > 0 || reps <= 0) best = std::min(best, (double)t);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ms) *ms = best;
  return MG_OK;
}
