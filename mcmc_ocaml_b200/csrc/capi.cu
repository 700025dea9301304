// capi.cu -- context management and memory helpers of the C ABI.
#include "common.cuh"

#include <cxxabi.h>

using namespace mg;

extern "C" int mg_abi_version(void) { return MG_ABI_VERSION; }

// Make sure the device's stream-ordered pool holds at least `nbytes` in one piece (allocate, free: it stays cached).
extern "C" int mg_ctx_reserve_pool(mg_ctx *ctx, int64_t nbytes) {
  if (!ctx) return MG_EINVAL;
  if (nbytes <= 0) return MG_OK;
  cudaSetDevice(ctx->device);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) != cudaSuccess) { cudaGetLastError(); return MG_OK; }
  uint64_t have = 0;
  cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &have);
  if ((int64_t)have >= nbytes) return MG_OK;
  void *p = nullptr;
  if (cudaMallocAsync(&p, (size_t)nbytes, ctx->stream) == cudaSuccess) {
    cudaFreeAsync(p, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  } else cudaGetLastError();                        // not enough memory for a reserve: the pool grows on demand instead
  return MG_OK;
}

extern "C" int mg_ctx_create(int device, uint64_t seed, mg_ctx **out) {
  if (!out) return MG_EINVAL;
  *out = nullptr;
  int ndev = 0;
  // No CPU fallback: without a CUDA device the library refuses to work.
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return MG_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return MG_ECUDA;
  mg_ctx *ctx = new mg_ctx;
  ctx->device = device; ctx->seed = seed; ctx->epoch = 0;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MG_ECUDA; }
  ctx->own_stream = true;
  // the second stream (draw-ahead of Nested, Stats passes of the resident sampler) exists from the start, so that no
  // path ever falls back to the legacy NULL stream (which would serialise with every blocking stream of the process)
  if (cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return MG_ECUDA; }
  if (cudaMalloc((void **)&ctx->d_devflag, 4 * sizeof(int)) != cudaSuccess || cudaMemset(ctx->d_devflag, 0, 4 * sizeof(int)) != cudaSuccess) {
    cudaStreamDestroy(ctx->aux); cudaStreamDestroy(ctx->stream); delete ctx; return MG_ECUDA;
  }
  cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // keep freed stream-ordered allocations cached in the pool between calls
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    // The pool is given one large piece up front (allocated and freed again: it stays cached).  The calls of this
    // library ask for temporaries of 1-2 GB in varying order; a pool that grew request by request answers some of them
    // by remapping its fragments -- measured on mg_evidence_lebesgue with host buffers at 1e7 x 20: 58-195 ms per call
    // without the reserve, 58.7-60 ms with 12 GB (profiles/r02_summary.md).  Default: 12 GB when the device has at
    // least 48 GB free; MCMC_GPU_POOL_RESERVE_GB overrides (0 = none); mg_ctx_trim_pool hands it back.
    {
      double gb = 12.0;
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < (48ull << 30)) gb = 0.0;
      if (const char *e = getenv("MCMC_GPU_POOL_RESERVE_GB")) gb = atof(e);
      if (gb > 0.0) mg_ctx_reserve_pool(ctx, (int64_t)(gb * 1073741824.0));
    }
  }
  *out = ctx;
  return MG_OK;
}

extern "C" void mg_ctx_destroy(mg_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (cudaEvent_t e : ctx->kt_ev) if (e) cudaEventDestroy(e);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
  if (ctx->d_devflag) cudaFree(ctx->d_devflag);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *mg_last_error(const mg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int mg_ctx_set_seed(mg_ctx *ctx, uint64_t seed) { if (!ctx) return MG_EINVAL; ctx->seed = seed; ctx->epoch = 0; return MG_OK; }
extern "C" int mg_ctx_set_epoch(mg_ctx *ctx, uint64_t epoch) { if (!ctx) return MG_EINVAL; ctx->epoch = epoch; return MG_OK; }
extern "C" uint64_t mg_ctx_get_epoch(const mg_ctx *ctx) { return ctx ? ctx->epoch : 0; }

extern "C" int mg_ctx_set_stream(mg_ctx *ctx, void *cuda_stream) {
  if (!ctx) return MG_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
  if (cuda_stream) ctx->stream = (cudaStream_t)cuda_stream;
  else {
    MG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return MG_OK;
}
extern "C" void *mg_ctx_get_stream(const mg_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int mg_ctx_sync(mg_ctx *ctx) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return poll_device_error(ctx);
}
// The library's temporaries come from the device's stream-ordered pool and stay cached there between calls
// (release threshold = unlimited, set in mg_ctx_create); this hands the cached memory back to the driver.
extern "C" int mg_ctx_trim_pool(mg_ctx *ctx) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaMemPool_t pool;
  MG_CUDA(ctx, cudaDeviceGetDefaultMemPool(&pool, ctx->device));
  MG_CUDA(ctx, cudaMemPoolTrimTo(pool, 0));
  return MG_OK;
}
extern "C" int mg_reset_counters(mg_ctx *ctx) { if (!ctx) return MG_EINVAL; ctx->naccept = ctx->nreject = 0; return MG_OK; }
extern "C" int mg_get_counters(mg_ctx *ctx, int64_t *naccept, int64_t *nreject) {
  if (!ctx) return MG_EINVAL;
  if (naccept) *naccept = ctx->naccept;
  if (nreject) *nreject = ctx->nreject;
  return MG_OK;
}
extern "C" int64_t mg_ctx_launch_count(const mg_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" double mg_ctx_last_kernel_ms(const mg_ctx *c) {
  mg_ctx *ctx = const_cast<mg_ctx *>(c);
  if (!ctx) return 0.0;
  if (ctx->ev_pending) {
    cudaSetDevice(ctx->device);
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess)
      ctx->last_kernel_ms = ms;
    ctx->ev_pending = false;
  }
  return ctx->last_kernel_ms;
}

// name / launches / mean duration of the dominant kernel of the last call (events around every launch)
extern "C" int mg_ctx_last_kernel_stats(mg_ctx *ctx, char *name, int64_t name_cap, int64_t *launches, double *mean_ms) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->kt_launches < 0) {
    double tot = 0.0;
    for (int i = 0; i < ctx->kt_used; ++i) {
      float ms = 0.f;
      if (cudaEventSynchronize(ctx->kt_ev[2 * i + 1]) == cudaSuccess &&
          cudaEventElapsedTime(&ms, ctx->kt_ev[2 * i], ctx->kt_ev[2 * i + 1]) == cudaSuccess) tot += ms;
    }
    ctx->kt_launches = ctx->kt_used;
    ctx->kt_mean_ms = ctx->kt_used ? tot / ctx->kt_used : 0.0;
    ctx->kt_name.clear();
    const char *mangled = nullptr;
    if (ctx->kt_func && cudaFuncGetName(&mangled, ctx->kt_func) == cudaSuccess && mangled) {
      int status = 0;
      char *dem = abi::__cxa_demangle(mangled, nullptr, nullptr, &status);
      ctx->kt_name = (status == 0 && dem) ? dem : mangled;
      free(dem);
      const size_t paren = ctx->kt_name.find('(');            // drop the parameter list
      if (paren != std::string::npos) ctx->kt_name.resize(paren);
      if (ctx->kt_name.rfind("void ", 0) == 0) ctx->kt_name.erase(0, 5);
    }
  }
  if (name && name_cap > 0) { strncpy(name, ctx->kt_name.c_str(), (size_t)name_cap - 1); name[name_cap - 1] = 0; }
  if (launches) *launches = ctx->kt_launches;
  if (mean_ms) *mean_ms = ctx->kt_mean_ms;
  return MG_OK;
}

extern "C" int mg_malloc_device(mg_ctx *ctx, int64_t nbytes, void **out) {
  if (!ctx || !out || nbytes < 0) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaMalloc(out, (size_t)(nbytes ? nbytes : 1)));
  return MG_OK;
}
extern "C" int mg_free_device(mg_ctx *ctx, void *p) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MG_CUDA(ctx, cudaFree(p));
  return MG_OK;
}
extern "C" int mg_malloc_pinned(mg_ctx *ctx, int64_t nbytes, void **out) {
  if (!ctx || !out || nbytes < 0) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaMallocHost(out, (size_t)(nbytes ? nbytes : 1)));
  return MG_OK;
}
extern "C" int mg_free_pinned(mg_ctx *ctx, void *p) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaFreeHost(p));
  return MG_OK;
}
extern "C" int mg_memcpy_h2d(mg_ctx *ctx, void *dst, const void *src, int64_t nbytes) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MG_OK;
}
extern "C" int mg_memcpy_d2h(mg_ctx *ctx, void *dst, const void *src, int64_t nbytes) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MG_OK;
}
extern "C" int mg_memcpy_d2d(mg_ctx *ctx, void *dst, const void *src, int64_t nbytes) {
  if (!ctx) return MG_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MG_OK;
}
