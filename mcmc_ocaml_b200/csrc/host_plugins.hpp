// host_plugins.hpp -- host-side validation / upload of plugin specs.
#pragma once
#include "common.cuh"
#include "models.cuh"

namespace mg {

bool user_kind_registered(int kind);  // jit.cu

inline int64_t logfn_expected_nparams(const mg_logfn *f) {
  const int64_t D = f->dim;
  switch (f->kind) {
    case MG_FN_ZERO: return 0;
    case MG_FN_CONST: return 1;
    case MG_FN_BOX_CLOSED: case MG_FN_BOX_OPEN: return 2 * D + 1;
    case MG_FN_GAUSS_DIAG: return 2 * D;
    case MG_FN_GAUSS_CORR: return D + D * (D + 1) / 2 + 1;
    case MG_FN_GAUSS_DATA: case MG_FN_CAUCHY_DATA: return -1;  // any
    case MG_FN_SHELL: return D + 2;
    case MG_FN_GAUSS_MIX:
      if (f->nparams < 1 || !f->params) return -2;
      return 1 + (int64_t)f->params[0] * D + D;
    default: return user_kind_registered(f->kind) ? -1 : -2;
  }
}

inline int validate_logfn(mg_ctx *ctx, const mg_logfn *f, int dim, const char *what) {
  if (!f) return set_err(ctx, MG_EINVAL, "%s: null plugin", what);
  if (f->dim != dim) return set_err(ctx, MG_EINVAL, "%s: plugin dim %d != %d", what, f->dim, dim);
  const int64_t need = logfn_expected_nparams(f);
  if (need == -2) return set_err(ctx, MG_EINVAL, "%s: unknown log-density kind %d", what, f->kind);
  if (need >= 0 && f->nparams != need)
    return set_err(ctx, MG_EINVAL, "%s: kind %d needs %lld params, got %lld", what, f->kind,
                   (long long)need, (long long)f->nparams);
  if ((f->kind == MG_FN_GAUSS_DATA || f->kind == MG_FN_CAUCHY_DATA) && dim != 2)
    return set_err(ctx, MG_EINVAL, "%s: data likelihoods are 2-D (mu, sigma)", what);
  if (f->nparams > 0 && !f->params) return set_err(ctx, MG_EINVAL, "%s: null params", what);
  return MG_OK;
}

inline int validate_proposal(mg_ctx *ctx, const mg_proposal *f, int dim) {
  if (!f) return set_err(ctx, MG_EINVAL, "jump_proposal: null plugin");
  if (f->dim != dim) return set_err(ctx, MG_EINVAL, "jump_proposal: plugin dim %d != %d", f->dim, dim);
  int64_t need;
  switch (f->kind) {
    case MG_PROP_BOX: need = dim; break;
    case MG_PROP_WRAP: need = 3 * (int64_t)dim; break;
    case MG_PROP_INDEP_GAUSS: need = 2 * (int64_t)dim; break;
    case MG_PROP_LEFT_BIASED: need = 1; if (dim != 1) return set_err(ctx, MG_EINVAL, "left-biased proposal is 1-D"); break;
    default: return set_err(ctx, MG_EINVAL, "jump_proposal: unknown kind %d", f->kind);
  }
  if (f->nparams != need || !f->params)
    return set_err(ctx, MG_EINVAL, "jump_proposal: kind %d needs %lld params, got %lld", f->kind,
                   (long long)need, (long long)f->nparams);
  return MG_OK;
}

// Device copies of plugin parameter blobs for the duration of one call.
struct DevLogFn {
  DevBuf<double> buf;
  DynFnParams params{};
  cudaError_t upload_from(const mg_logfn *f, cudaStream_t s) {
    cudaError_t e = upload(buf, f->params, (size_t)f->nparams, s);
    params.kind = f->kind; params.dim = f->dim; params.scale = f->scale;
    params.p = buf.get(); params.np = f->nparams;
    return e;
  }
};
struct DevProposal {
  DevBuf<double> buf;
  DynPropParams params{};
  cudaError_t upload_from(const mg_proposal *f, cudaStream_t s) {
    cudaError_t e = upload(buf, f->params, (size_t)f->nparams, s);
    params.kind = f->kind; params.dim = f->dim; params.p = buf.get(); params.np = f->nparams;
    return e;
  }
};

}  // namespace mg
