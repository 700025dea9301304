// host_plugins.hpp -- host-side validation / upload of plugin specs.
#pragma once
#include <cmath>
#include <vector>
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
    case MG_PROP_ONE_SIDED: need = 2; if (dim != 1) return set_err(ctx, MG_EINVAL, "one-sided proposal is 1-D"); break;
    case MG_PROP_MIXTURE: {  // K, then (weight, kind, nparams, params...) per component
      if (f->nparams < 1 || !f->params) return set_err(ctx, MG_EINVAL, "combine_jump_proposals: no components");
      const int K = (int)f->params[0];
      int64_t k = 1;
      if (K < 1 || K > 64) return set_err(ctx, MG_EINVAL, "combine_jump_proposals: need 1..64 components");
      for (int c = 0; c < K; ++c) {
        if (k + 3 > f->nparams) return set_err(ctx, MG_EINVAL, "combine_jump_proposals: truncated parameter block");
        mg_proposal comp{(int32_t)f->params[k + 1], dim, f->params + k + 3, (int64_t)f->params[k + 2]};
        if (comp.kind == MG_PROP_MIXTURE || comp.nparams < 0 || k + 3 + comp.nparams > f->nparams || !(f->params[k] > 0.0))
          return set_err(ctx, MG_EINVAL, "combine_jump_proposals: bad component %d", c);
        const int rc = validate_proposal(ctx, &comp, dim);
        if (rc) return rc;
        k += 3 + comp.nparams;
      }
      need = k; break;
    }
    case MG_PROP_DE: {       // mode_hopping_frac, M, samples[M][D]
      if (f->nparams < 2 || !f->params) return set_err(ctx, MG_EINVAL, "differential_evolution_proposal: no samples");
      const double M = f->params[1];
      if (!(M >= 2.0) || M != (double)(int64_t)M || M > 1e12) return set_err(ctx, MG_EINVAL, "differential_evolution_proposal: need at least two samples");
      need = 2 + (int64_t)M * dim; break;
    }
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
  // Values that depend on the parameters only are evaluated here, once, and appended after the nparams user values
  // (models.cuh reads them at p + np): log sigma of SHELL, log sigma_i of GAUSS_DIAG and GAUSS_MIX.
  cudaError_t upload_from(const mg_logfn *f, cudaStream_t s) {
    std::vector<double> blob(f->params, f->params + f->nparams);
    const int d = f->dim;
    if (f->kind == MG_FN_SHELL) blob.push_back(std::log(f->params[d + 1]));
    else if (f->kind == MG_FN_GAUSS_DIAG) for (int i = 0; i < d; ++i) blob.push_back(std::log(f->params[d + i]));
    else if (f->kind == MG_FN_GAUSS_MIX) {
      const int K = (int)f->params[0];
      for (int i = 0; i < d; ++i) blob.push_back(std::log(f->params[1 + K * d + i]));
    }
    cudaError_t e = upload(buf, blob.data(), blob.size(), s);   // pageable source: staged before the call returns
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
