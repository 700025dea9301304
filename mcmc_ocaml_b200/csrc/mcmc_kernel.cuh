// mcmc_kernel.cuh -- host side (launch) of the fused Metropolis-Hastings
// ensemble kernel; the kernel itself is in mcmc_kernel_dev.cuh.
// Instantiated per plugin combination in mcmc_static.cu (compile-time D,
// parameters staged in shared memory) and mcmc_dyn.cu (registered kinds).
#pragma once
#include "common.cuh"
#include "host_plugins.hpp"
#include "mcmc_kernel_dev.cuh"

namespace mg {

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
static void fill_common(A &a, const mg_mcmc_cfg *cfg, CallKey key, uint64_t t0, int record_first, double *d_state, double *d_samples,
                        int32_t *d_accept) {
  a.d = cfg->dim; a.pad = 0; a.C = cfg->nchains; a.chain_offset = cfg->chain_offset;
  a.nbin = cfg->nbin; a.nskip = cfg->nskip; a.n = cfg->n; a.key = key;
  a.t0 = t0; a.record_first = record_first; a.pad2 = 0;
  a.rk = make_round_keys(key);
  a.state = d_state; a.samples = d_samples; a.accept = d_accept;
}

}  // namespace mg
