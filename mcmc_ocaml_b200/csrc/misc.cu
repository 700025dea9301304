// misc.cu -- small entry points: plugin evaluation on a batch of points,
// Mcmc.remove_repeat_samples, user-plugin registration.
#include "common.cuh"
#include "host_plugins.hpp"
#include "models.cuh"
#include "scan.cuh"

namespace mg {

int jit_logfn_eval(mg_ctx *ctx, const DynFnParams &f, const double *d_x, int64_t M, double *d_out);  // jit.cu

template <int DMAX>
__global__ void logfn_eval_kernel(DynFnParams f, const double *__restrict__ x, int64_t M, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double v[DMAX];
#pragma unroll (DMAX <= 16 ? DMAX : 1)
  for (int d = 0; d < DMAX; ++d) v[d] = (d < f.dim) ? x[i * f.dim + d] : 0.0;
  out[i] = DynFn::eval<DMAX>(f, nullptr, v, f.dim);
}

// keep[i] = 1 iff i = 0 or value_i <> value_{i-1}  (mcmc.ml:74-81, eql = (=))
__global__ void repeat_flags_kernel(const double *__restrict__ rows, int64_t n, int D, int32_t *__restrict__ keep) {
  const int F = D + 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int k = 1;
    if (i > 0) {
      bool eq = true;
      for (int d = 0; d < D && eq; ++d) eq = (rows[i * F + d] == rows[(i - 1) * F + d]);
      k = eq ? 0 : 1;
    }
    keep[i] = k;
  }
}
__global__ void repeat_gather_kernel(const double *__restrict__ rows, int64_t n, int F, const int32_t *__restrict__ keep,
                                     const int32_t *__restrict__ rank, double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (keep[i]) for (int f = 0; f < F; ++f) out[(int64_t)rank[i] * F + f] = rows[i * F + f];
}

}  // namespace mg

using namespace mg;

extern "C" int mg_logfn_eval(mg_ctx *ctx, const mg_logfn *fn, const double *x, int64_t M, double *out) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, fn && x && out && M >= 0, "logfn_eval: bad arguments");
  MG_REQUIRE(ctx, fn->dim >= 1 && fn->dim <= 64, "logfn_eval: dim must be in 1..64");
  int rc = validate_logfn(ctx, fn, fn->dim, "log-density");
  if (rc) return rc;
  if (M == 0) return MG_OK;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevLogFn df;
  DevBuf<double> d_x, d_out;
  MG_CUDA(ctx, df.upload_from(fn, s));
  MG_CUDA(ctx, upload(d_x, x, (size_t)M * fn->dim, s));
  MG_CUDA(ctx, d_out.alloc(M, s));
  const unsigned grid = (unsigned)((M + 127) / 128);
  const int D = fn->dim;
  if (fn->kind >= MG_FN_USER) {
    if ((rc = jit_logfn_eval(ctx, df.params, d_x.get(), M, d_out.get()))) return rc;
  } else if (D <= 2) logfn_eval_kernel<2><<<grid, 128, 0, s>>>(df.params, d_x.get(), M, d_out.get());
  else if (D <= 4) logfn_eval_kernel<4><<<grid, 128, 0, s>>>(df.params, d_x.get(), M, d_out.get());
  else if (D <= 8) logfn_eval_kernel<8><<<grid, 128, 0, s>>>(df.params, d_x.get(), M, d_out.get());
  else if (D <= 16) logfn_eval_kernel<16><<<grid, 128, 0, s>>>(df.params, d_x.get(), M, d_out.get());
  else if (D <= 32) logfn_eval_kernel<32><<<grid, 128, 0, s>>>(df.params, d_x.get(), M, d_out.get());
  else logfn_eval_kernel<64><<<grid, 128, 0, s>>>(df.params, d_x.get(), M, d_out.get());
  MG_CHECK_LAUNCH(ctx);
  MG_CUDA(ctx, cudaMemcpyAsync(out, d_out.get(), sizeof(double) * M, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  return MG_OK;
}

extern "C" int mg_remove_repeat_samples(mg_ctx *ctx, const double *rows, int64_t n, int32_t dim, double *out,
                                        int64_t *nkept) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, rows && out && nkept && n >= 0 && dim >= 1, "remove_repeat_samples: bad arguments");
  *nkept = 0;
  if (n == 0) return MG_OK;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int F = dim + 2;
  DevBuf<double> d_rows, d_out;
  DevBuf<int32_t> keep, rank, tmp, total;
  MG_CUDA(ctx, upload(d_rows, rows, (size_t)n * F, s));
  MG_CUDA(ctx, d_out.alloc((size_t)n * F, s));
  MG_CUDA(ctx, keep.alloc(n, s)); MG_CUDA(ctx, rank.alloc(n, s));
  MG_CUDA(ctx, tmp.alloc((size_t)scan_tmp_elems(n, 1) + 1, s)); MG_CUDA(ctx, total.alloc(1, s));
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
  repeat_flags_kernel<<<grid, 256, 0, s>>>(d_rows.get(), n, dim, keep.get());
  MG_CHECK_LAUNCH(ctx);
  int rc = exclusive_scan_i32(ctx, keep.get(), rank.get(), n, 1, tmp.get(), total.get());
  if (rc) return rc;
  repeat_gather_kernel<<<grid, 256, 0, s>>>(d_rows.get(), n, F, keep.get(), rank.get(), d_out.get());
  MG_CHECK_LAUNCH(ctx);
  int32_t K = 0;
  MG_CUDA(ctx, cudaMemcpyAsync(&K, total.get(), 4, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  MG_CUDA(ctx, cudaMemcpyAsync(out, d_out.get(), sizeof(double) * (size_t)K * F, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  *nkept = K;
  return MG_OK;
}

// ---- diagnostic: accept-test prefilter against the plain float64 comparison --------------------------------
namespace mg {
__global__ void accept_test_kernel(const double *__restrict__ u, const double *__restrict__ delta, int64_t n,
                                   uint8_t *__restrict__ fast, uint8_t *__restrict__ exact) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    fast[i] = log_u_less_than(u[i], delta[i]) ? 1 : 0;
    exact[i] = (log(u[i]) < delta[i]) ? 1 : 0;
  }
}
}  // namespace mg

extern "C" int mg_debug_accept_test(mg_ctx *ctx, const double *u, const double *delta, int64_t n, uint8_t *out_fast,
                                    uint8_t *out_exact) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, u && delta && out_fast && out_exact && n >= 0, "debug_accept_test: bad arguments");
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf<double> du, dd;
  DevBuf<uint8_t> df, de;
  MG_CUDA(ctx, upload(du, u, (size_t)n, s));
  MG_CUDA(ctx, upload(dd, delta, (size_t)n, s));
  MG_CUDA(ctx, df.alloc((size_t)n + 1, s));
  MG_CUDA(ctx, de.alloc((size_t)n + 1, s));
  if (n) {
    accept_test_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, s>>>(du.get(), dd.get(), n, df.get(), de.get());
    MG_CHECK_LAUNCH(ctx);
    MG_CUDA(ctx, cudaMemcpyAsync(out_fast, df.get(), (size_t)n, cudaMemcpyDeviceToHost, s));
    MG_CUDA(ctx, cudaMemcpyAsync(out_exact, de.get(), (size_t)n, cudaMemcpyDeviceToHost, s));
  }
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  return MG_OK;
}
