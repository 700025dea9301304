// nested_kernel_dev.cuh -- the parts of Nested.nested_evidence (nested.ml:50-74,122-132) that evaluate the
// log-likelihood and the log-prior, as device bodies without host includes: they are compiled into libmcmcgpu.so and,
// when a run uses a user-registered log-density, once more at run time by NVRTC with that function inlined (jit.cu).
#pragma once
#include "models.cuh"
#include "rng.cuh"

namespace mg {

struct NestArgs {
  DynFnParams like, prior;
  const double *live_x, *live_ll, *live_lp;   // sorted ascending in ll
  double *fresh_x, *fresh_ll, *fresh_lp;      // [K][D], [K], [K]
  const double *plo, *phi;                    // prior box
  CallKey key;
  int64_t R;                                  // replacements done so far
  double threshold, mode_hop, de_sigma;
  int32_t D, nlive, K, nmcmc;
  int *fail;
};

// draw_prior (nested_test.ml:34-35 style: per-dimension Stats.draw_uniform) + evaluation (:126-130)
template <int DMAX>
__device__ __forceinline__ void nest_init_body(const NestArgs &a, double *__restrict__ x_out, double *__restrict__ ll_out,
                                               double *__restrict__ lp_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nlive) return;
  Rng r(a.key, P_NEST_INIT, (uint64_t)i, 0);
  double x[DMAX];
#pragma unroll (DMAX <= 16 ? DMAX : 1)
  for (int d = 0; d < DMAX; ++d) x[d] = (d < a.D) ? draw_uniform(r, __ldg(a.plo + d), __ldg(a.phi + d)) : 0.0;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
  for (int d = 0; d < DMAX; ++d)
    if (d < a.D) x_out[(int64_t)i * a.D + d] = x[d];
  ll_out[i] = DynFn::eval<DMAX>(a.like, nullptr, x, a.D);
  lp_out[i] = DynFn::eval<DMAX>(a.prior, nullptr, x, a.D);
}


struct NestProp {
  int32_t *i0, *j0;   // [S][K] live-set rows x, y of the proposal  (mcmc.ml:201-202)
  double *ds;         // [S][K] scale d                             (:209-213)
  double *u;          // [S][K] Random.float 1.0 of the accept test (mcmc.ml:47)
};


// draw_new_live_point (nested.ml:50-74) for one chain per thread, plain loads: the form that is compiled at run time
// for user plugins.  It reads the same pre-drawn proposals (nest_propose_kernel) and does the same arithmetic in the
// same order as the pipelined nest_replace_kernel of nested.cu, so both give the same chains
// (MCMC_GPU_NEST_SIMPLE=1 runs this form for the built-in plugins; tests/test_nested_gpu.py compares the two).
template <int DMAX>
__device__ __forceinline__ void nest_replace_simple_body(const NestArgs &a, const NestProp &p, int s0, int s1, int first, int last,
                                                         double *chain_x, double *chain_cl) {
  const int K = a.K;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K) return;
  const int D = a.D;
  const double thr = a.threshold;
  auto mcmc_logl = [&](const double (&pt)[DMAX]) {           // nested.ml:54-59
    const double l = DynFn::eval<DMAX>(a.like, nullptr, pt, D);
    const double pr = DynFn::eval<DMAX>(a.prior, nullptr, pt, D);
    return (l >= thr) ? pr : neg_inf();
  };
  double x[DMAX], y[DMAX];
  double cl;
  if (first) {
    Rng rs(a.key, P_NEST_START, (uint64_t)(a.R + j), 0);
    const int start = (a.K - 1) + (int)rs.below((uint64_t)(a.nlive - a.K + 1));   // :63 (above the common threshold)
    for (int d = 0; d < DMAX; ++d) x[d] = (d < D) ? a.live_x[(int64_t)start * D + d] : 0.0;
    cl = mcmc_logl(x);
  } else {
    for (int d = 0; d < DMAX; ++d) x[d] = (d < D) ? chain_x[(int64_t)d * K + j] : 0.0;
    cl = chain_cl[j];
  }
  const double cp = 0.0;                                     // mcmc_logp, :60
  const int S = s1 - s0;
  for (int s = 0; s < S; ++s) {                              // :65-67
    const int64_t q = (int64_t)s * K + j;
    const double ds = p.ds[q], u_cur = p.u[q];
    const double *rx = a.live_x + (int64_t)p.i0[q] * D, *ry = a.live_x + (int64_t)p.j0[q] * D;
    for (int d = 0; d < DMAX; ++d) {
      const double delta = (d < D) ? ds * (ry[d] - rx[d]) : 0.0;   // mcmc.ml:214-216
      y[d] = (d < D) ? x[d] + delta : 0.0;
    }
    const double start_log_post = cl + cp;
    const double proposed_like = mcmc_logl(y);
    const double proposed_log_posterior = proposed_like + 0.0;
    const double log_accept_prob = proposed_log_posterior - start_log_post + 0.0 - 0.0;
    if (log_u_less_than(u_cur, log_accept_prob)) {
      for (int d = 0; d < DMAX; ++d) x[d] = y[d];
      cl = proposed_like;
    }
  }
  if (!last) {
    for (int d = 0; d < DMAX; ++d)
      if (d < D) chain_x[(int64_t)d * K + j] = x[d];
    chain_cl[j] = cl;
    return;
  }
  const double nl = DynFn::eval<DMAX>(a.like, nullptr, x, D);      // :68-69
  const double np = DynFn::eval<DMAX>(a.prior, nullptr, x, D);
  if (!(nl >= thr)) *a.fail = 1;                                   // :70-72
  for (int d = 0; d < DMAX; ++d)
    if (d < D) a.fresh_x[(int64_t)j * D + d] = x[d];
  a.fresh_ll[j] = nl; a.fresh_lp[j] = np;
}

}  // namespace mg
