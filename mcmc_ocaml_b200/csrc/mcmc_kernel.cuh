// mcmc_kernel.cuh -- the fused Metropolis-Hastings ensemble kernel template.
// Instantiated per plugin combination in mcmc_static.cu (compile-time D,
// parameters in the constant bank) and mcmc_dyn.cu (registered kinds).
#pragma once
#include "common.cuh"
#include "host_plugins.hpp"
#include "models.cuh"

namespace mg {

constexpr int MH_BLOCK = 64;  // 65,536 chains -> 1024 CTAs = 6.9 per SM on 148 SMs

template <class Like, class Prior, class Prop, int D>
struct MhArgs {
  typename Like::Params like;
  typename Prior::Params prior;
  typename Prop::Params prop;
  int32_t d, pad;
  int64_t C;
  uint64_t chain_offset;
  int64_t nbin, nskip, n;
  CallKey key;
  double *state;    // [D+2][C] in/out
  double *samples;  // [n][D+2][C] or null
  int32_t *accept;  // [C] accumulated, or null
};

// mcmc.ml:37-56 make_mcmc_sampler: one step.  Returns 1 on acceptance.
template <class Like, class Prior, class Prop, int D>
__device__ __forceinline__ int mh_step(const MhArgs<Like, Prior, Prop, D> &a, uint64_t g, uint64_t t,
                                       double (&x)[D], double &ll, double &lp) {
  Rng r(a.key, P_MH, g, t);
  const double start_log_post = ll + lp;
  double y[D];
  Prop::template propose<D>(a.prop, r, x, y, a.d);
  const double proposed_like = Like::template eval<D>(a.like, y, a.d);
  const double proposed_prior = Prior::template eval<D>(a.prior, y, a.d);
  const double proposed_log_posterior = proposed_like + proposed_prior;
  double log_accept_prob = proposed_log_posterior - start_log_post;
  if (!Prop::kSymmetric) {
    const double log_forward_jump = Prop::template log_q<D>(a.prop, x, y, a.d);
    const double log_backward_jump = Prop::template log_q<D>(a.prop, y, x, a.d);
    log_accept_prob = log_accept_prob + log_backward_jump - log_forward_jump;
  }
  // log (Random.float 1.0) < log_accept_prob, strict (mcmc.ml:47).  NaN rejects.
  const bool acc = log(r.uniform()) < log_accept_prob;
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = acc ? y[i] : x[i];
  ll = acc ? proposed_like : ll;
  lp = acc ? proposed_prior : lp;
  return acc ? 1 : 0;
}

template <class Like, class Prior, class Prop, int D>
__global__ void __launch_bounds__(MH_BLOCK)
mh_ensemble_kernel(const __grid_constant__ MhArgs<Like, Prior, Prop, D> a) {
  const int64_t c = (int64_t)blockIdx.x * MH_BLOCK + threadIdx.x;
  if (c >= a.C) return;
  const uint64_t g = a.chain_offset + (uint64_t)c;
  const int64_t C = a.C;
  const int F = a.d + 2;
  double x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = (i < a.d) ? a.state[(int64_t)i * C + c] : 0.0;
  // mcmc.ml:59-61: the start point is evaluated, not trusted
  double ll = Like::template eval<D>(a.like, x, a.d);
  double lp = Prior::template eval<D>(a.prior, x, a.d);
  int nacc = 0;
  uint64_t t = 0;
  for (int64_t i = 0; i < a.nbin; ++i, ++t) nacc += mh_step<Like, Prior, Prop, D>(a, g, t, x, ll, lp);  // :63-65
  double *out = a.samples ? a.samples + c : nullptr;
  const int64_t sample_stride = (int64_t)F * C;
  auto record = [&]() {
    if (out) {
#pragma unroll
      for (int i = 0; i < D; ++i)
        if (i < a.d) __stcs(out + (int64_t)i * C, x[i]);
      __stcs(out + (int64_t)a.d * C, ll);
      __stcs(out + (int64_t)(a.d + 1) * C, lp);
      out += sample_stride;
    }
  };
  if (a.n > 0) record();  // :66 slot 0 = state after burn-in
  for (int64_t s = 1; s < a.n; ++s) {  // :67-71
    for (int64_t k = 0; k < a.nskip; ++k, ++t) nacc += mh_step<Like, Prior, Prop, D>(a, g, t, x, ll, lp);
    record();
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
    if (i < a.d) a.state[(int64_t)i * C + c] = x[i];
  a.state[(int64_t)a.d * C + c] = ll;
  a.state[(int64_t)(a.d + 1) * C + c] = lp;
  if (a.accept) a.accept[c] += nacc;
}

template <class Like, class Prior, class Prop, int D>
static int launch_mh(mg_ctx *ctx, const MhArgs<Like, Prior, Prop, D> &a) {
  const int64_t grid = (a.C + MH_BLOCK - 1) / MH_BLOCK;
  time_begin(ctx);
  mh_ensemble_kernel<Like, Prior, Prop, D><<<(unsigned)grid, MH_BLOCK, 0, ctx->stream>>>(a);
  MG_CHECK_LAUNCH(ctx);
  time_end(ctx);
  return MG_OK;
}

template <class A>
static void fill_common(A &a, const mg_mcmc_cfg *cfg, CallKey key, double *d_state, double *d_samples,
                        int32_t *d_accept) {
  a.d = cfg->dim; a.pad = 0; a.C = cfg->nchains; a.chain_offset = cfg->chain_offset;
  a.nbin = cfg->nbin; a.nskip = cfg->nskip; a.n = cfg->n; a.key = key;
  a.state = d_state; a.samples = d_samples; a.accept = d_accept;
}

}  // namespace mg
