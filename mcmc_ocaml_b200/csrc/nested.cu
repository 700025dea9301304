// nested.cu -- Nested.nested_evidence (nested.ml:122-146) with batched
// constrained live-point replacement.
//
// The reference retires ONE live point per iteration and replaces it with the
// end point of an nmcmc-step constrained Metropolis-Hastings chain whose
// proposal is differential evolution over the current live set
// (nested.ml:50-74, mcmc.ml:198-218): nmcmc sequential likelihood calls per
// retired point.  Here the K lowest points are retired together
// (cfg.batch = K; K = 1 is the reference's schedule): K independent
// replacement chains run as one kernel, one thread per chain, all constrained
// to ll >= the K-th lowest log-likelihood; the j-th point of a batch leaves
// nlive - j points above it, so its shrinkage factor is 1 - 1/(nlive - j)
// (with j = 0 always, K = 1 reproduces nested.ml:131,139 operation for
// operation).  The live set ([nlive][D] rows, 12.8 MB at BASELINE config 4)
// stays L2-resident; the DE proposal gathers two rows per step.
#include "common.cuh"
#include "host_plugins.hpp"
#include "models.cuh"
#include "radix_sort.cuh"
#include "reduce_sum.cuh"

#include <algorithm>
#include <cmath>

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
__global__ void nest_init_kernel(NestArgs a, double *__restrict__ x_out, double *__restrict__ ll_out,
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

// draw_new_live_point (nested.ml:50-74), one thread per replacement chain
template <int DMAX>
__global__ void nest_replace_kernel(NestArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.K) return;
  const uint64_t rid = (uint64_t)(a.R + j);
  Rng rs(a.key, P_NEST_START, rid, 0);
  // livepts.(Random.int nlive) (:63); with K > 1 the start must satisfy the common threshold
  const int start = (a.K - 1) + (int)rs.below((uint64_t)(a.nlive - a.K + 1));
  double x[DMAX], y[DMAX];
#pragma unroll (DMAX <= 16 ? DMAX : 1)
  for (int d = 0; d < DMAX; ++d) x[d] = (d < a.D) ? a.live_x[(int64_t)start * a.D + d] : 0.0;
  const double thr = a.threshold;
  auto mcmc_logl = [&](const double (&pt)[DMAX]) {           // :54-59
    const double l = DynFn::eval<DMAX>(a.like, nullptr, pt, a.D);
    return (l >= thr) ? DynFn::eval<DMAX>(a.prior, nullptr, pt, a.D) : neg_inf();
  };
  double cl = mcmc_logl(x);
  const double cp = 0.0;                                     // mcmc_logp, :60
  for (int s = 0; s < a.nmcmc; ++s) {                        // :65-67
    Rng r(a.key, P_NEST_MCMC, rid, (uint64_t)s);
    // differential_evolution_proposal (mcmc.ml:198-218)
    const uint64_t n = (uint64_t)a.nlive;
    const uint64_t i0 = r.below(n);
    uint64_t j0;
    do { j0 = r.below(n); } while (j0 == i0);
    double dscale;
    if (a.mode_hop != 0.0 && r.uniform() < a.mode_hop) dscale = 1.0;
    else dscale = draw_gaussian(r, 0.0, a.de_sigma);
    const double *px = a.live_x + (int64_t)i0 * a.D, *py = a.live_x + (int64_t)j0 * a.D;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
    for (int d = 0; d < DMAX; ++d) y[d] = (d < a.D) ? x[d] + dscale * (__ldg(py + d) - __ldg(px + d)) : 0.0;
    // make_mcmc_sampler (mcmc.ml:37-56) with the closures of :54-61
    const double start_log_post = cl + cp;
    const double proposed_like = mcmc_logl(y);
    const double proposed_log_posterior = proposed_like + 0.0;
    const double log_accept_prob = proposed_log_posterior - start_log_post + 0.0 - 0.0;
    if (log_u_less_than(r.uniform(), log_accept_prob)) {
#pragma unroll (DMAX <= 16 ? DMAX : 1)
      for (int d = 0; d < DMAX; ++d) x[d] = y[d];
      cl = proposed_like;
    }
  }
  const double nl = DynFn::eval<DMAX>(a.like, nullptr, x, a.D);   // :68-69
  const double np = DynFn::eval<DMAX>(a.prior, nullptr, x, a.D);
  if (!(nl >= thr)) *a.fail = 1;                                   // :70-72
#pragma unroll (DMAX <= 16 ? DMAX : 1)
  for (int d = 0; d < DMAX; ++d)
    if (d < a.D) a.fresh_x[(int64_t)j * a.D + d] = x[d];
  a.fresh_ll[j] = nl; a.fresh_lp[j] = np;
}

__global__ void nest_keys_kernel(const double *__restrict__ ll, int n, uint64_t *__restrict__ keys, int32_t *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = f64_to_ordered(ll[i]); vals[i] = i; }
}
__global__ void nest_gather_kernel(const double *__restrict__ x, const double *__restrict__ ll,
                                   const double *__restrict__ lp, const int32_t *__restrict__ order, int n, int D,
                                   double *__restrict__ xo, double *__restrict__ llo, double *__restrict__ lpo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = order[i];
  for (int d = 0; d < D; ++d) xo[(int64_t)i * D + d] = x[(int64_t)s * D + d];
  llo[i] = ll[s]; lpo[i] = lp[s];
}

// Shrinkage schedule shared with the oracle (oracle.cpp `Shrink`): host libm.
struct Shrink {
  int nlive, K;
  std::vector<double> s, lvf;  // partial sums of log1p(-1/(nlive-j)); log(1/(nlive-j))
  double S;
  Shrink(int nlive_, int K_) : nlive(nlive_), K(K_), s(K_ + 1, 0.0), lvf(K_, 0.0) {
    for (int j = 0; j < K; ++j) {
      s[j + 1] = s[j] + std::log1p(-(1.0 / (double)(nlive - j)));
      lvf[j] = std::log(1.0 / (double)(nlive - j));
    }
    S = s[K];
  }
};

// stats.ml:240-248 on the host (glibc), used by the running tracker
static inline double h_log_sum_logs(double a, double b) {
  if (a == -INFINITY && b == -INFINITY) return -INFINITY;
  if (b > a) std::swap(a, b);
  return a + std::log1p(std::exp(b - a));
}

// evidence_error_and_weights (nested.ml:81-120), one thread per point.
// tab: [0..K) s_j, [K..2K) log vol_fraction_j, [2K] S
__global__ void nest_weights_kernel(const double *__restrict__ ll, int64_t n, int64_t ilive, int K,
                                    const double *__restrict__ tab, double log_dv_live, double *__restrict__ wts,
                                    double *__restrict__ dlow, double *__restrict__ dhigh) {
  const double log_half = -0.69314718055994530942;
  const double S = tab[2 * K];
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    auto log_dv = [&](int64_t i) { return tab[K + (int)(i % K)] + ((double)(i / K) * S + tab[(int)(i % K)]); };
    double w = neg_inf();
    if (k < ilive) {
      if (k >= 1) w = log_sum_logs(w, log_half + (log_dv(k - 1) + ll[k]));      // dhigh of iteration k-1
      const double dl = log_dv(k) + ll[k];
      w = log_sum_logs(w, log_half + dl);                                       // dlow of iteration k
      dlow[k] = dl; dhigh[k] = log_dv(k) + ll[k + 1];
      if (k == ilive - 1) w = log_sum_logs(w, log_half + (log_dv_live + ll[k])); // tail iteration i = ilive
    } else {
      if (k == ilive && ilive >= 1) w = log_sum_logs(w, log_half + (log_dv(ilive - 1) + ll[k]));
      const double t = log_half + (log_dv_live + ll[k]);
      if (k >= 1) w = log_sum_logs(w, t);                 // tail iteration i = k   (skipped for i = 0)
      if (k + 1 <= n - 1) w = log_sum_logs(w, t);         // tail iteration i = k+1
      // tail iteration i = k contributes dlow = dv + ll[k-1], dhigh = dv + ll[k]
      if (k >= 1) { dlow[k] = log_dv_live + ll[k - 1]; dhigh[k] = log_dv_live + ll[k]; }
      else { dlow[k] = neg_inf(); dhigh[k] = neg_inf(); }
    }
    wts[k] = w;
  }
}

__global__ void nest_normalise_kernel(double *__restrict__ wts, int64_t n, double log_ev) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    wts[k] = wts[k] - log_ev;
}

__global__ void max_kernel(const double *__restrict__ x, int64_t n, unsigned long long *__restrict__ out) {
  unsigned long long m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = f64_to_ordered(x[i]);
    m = k > m ? k : m;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, off); m = o > m ? o : m; }
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

static inline double ordered_to_f64(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  double d; memcpy(&d, &b, 8); return d;
}

// log sum_i exp(x_i) over a device array (max + compensated sum)
static int device_logsumexp(mg_ctx *ctx, const double *d_x, int64_t n, double *out) {
  cudaStream_t s = ctx->stream;
  DevBuf<unsigned long long> d_m;
  MG_CUDA(ctx, d_m.alloc(1, s));
  MG_CUDA(ctx, cudaMemsetAsync(d_m.get(), 0, 8, s));
  max_kernel<<<egrid(ctx, n), EB, 0, s>>>(d_x, n, d_m.get());
  MG_CHECK_LAUNCH(ctx);
  unsigned long long hm = 0;
  MG_CUDA(ctx, cudaMemcpyAsync(&hm, d_m.get(), 8, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  const double m = ordered_to_f64(hm);
  if (m == -INFINITY) { *out = -INFINITY; return MG_OK; }
  double sum = 0.0;
  int rc = reduce_sum(ctx, n, [d_x, m] __device__(int64_t i) { return exp(d_x[i] - m); }, &sum);
  if (rc) return rc;
  *out = m + std::log(sum);
  return MG_OK;
}

static int weights_dev(mg_ctx *ctx, const double *d_ll, int64_t n, int nlive, int K, double ll_last_retired_unused,
                       double *log_ev, double *log_dev, double *d_wts) {
  (void)ll_last_retired_unused;
  cudaStream_t s = ctx->stream;
  Shrink sh(nlive, K);
  std::vector<double> tab(2 * K + 1);
  for (int j = 0; j < K; ++j) { tab[j] = sh.s[j]; tab[K + j] = sh.lvf[j]; }
  tab[2 * K] = sh.S;
  const int64_t ilive = n - nlive;
  // nested.ml:97: log_vol_fraction + (ilive-1) * log_reduction_frac
  const int64_t il1 = ilive - 1;
  double log_x_il1;
  if (il1 >= 0) log_x_il1 = (double)(il1 / K) * sh.S + sh.s[il1 % K];
  else log_x_il1 = (double)il1 * sh.S;  // ilive = 0: (float_of_int (-1)) *. log_reduction_frac (K = 1 form)
  const double log_dv_live = std::log(1.0 / (double)nlive) + log_x_il1;
  DevBuf<double> d_tab, d_dlow, d_dhigh;
  MG_CUDA(ctx, upload(d_tab, tab.data(), tab.size(), s));
  MG_CUDA(ctx, d_dlow.alloc(n, s));
  MG_CUDA(ctx, d_dhigh.alloc(n, s));
  nest_weights_kernel<<<egrid(ctx, n), EB, 0, s>>>(d_ll, n, ilive, K, d_tab.get(), log_dv_live, d_wts, d_dlow.get(), d_dhigh.get());
  MG_CHECK_LAUNCH(ctx);
  double low, high;
  int rc;
  if ((rc = device_logsumexp(ctx, d_dlow.get(), n, &low))) return rc;
  if ((rc = device_logsumexp(ctx, d_dhigh.get(), n, &high))) return rc;
  const double log_half = -0.69314718055994530942;
  *log_ev = log_half + h_log_sum_logs(low, high);                 // :114
  *log_dev = high + std::log1p(-std::exp(low - high));            // :115
  nest_normalise_kernel<<<egrid(ctx, n), EB, 0, s>>>(d_wts, n, *log_ev);
  MG_CHECK_LAUNCH(ctx);
  return MG_OK;
}

}  // namespace mg

using namespace mg;

// nested.ml:152-178 on the host: n is small (the reference asks for ~100 draws) and the running sums are a
// sequential float64 scan in the reference's order.
extern "C" int mg_nested_posterior_indices(mg_ctx *ctx, const double *logw, int64_t npts, int64_t n, int64_t *out_idx) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, logw && out_idx && npts >= 1 && n >= 0, "posterior_samples: bad arguments");
  const CallKey key = next_key(ctx);
  std::vector<double> sw((size_t)npts);
  sw[0] = std::exp(logw[0]);
  for (int64_t i = 1; i < npts; ++i) sw[i] = std::exp(logw[i]) + sw[i - 1];      // :171-173
  for (int64_t k = 0; k < n; ++k) {
    uint32_t w[4];
    const uint64_t g = (uint64_t)k;
    philox4x32_10(0u, 0u, (uint32_t)g, (uint32_t)((g >> 32) & 0xFFFFu) | ((uint32_t)P_POST << 16), key.k0, key.k1, w);
    uint64_t bits = (0x3FFull << 52) | ((((uint64_t)w[0] << 32) | w[1]) & 0xFFFFFFFFFFFFFull);
    double m; memcpy(&m, &bits, 8);
    const double x = m - 1.0;                                                     // Random.float 1.0
    int64_t idx;
    if (x <= sw[0]) idx = 0;                                                      // weight_binary_search_index :152-165
    else { int64_t lo = 0, hi = npts - 1; while (hi - lo > 1) { const int64_t mid = (lo + hi) / 2; if (x <= sw[mid]) hi = mid; else lo = mid; } idx = hi; }
    out_idx[k] = idx;
  }
  return MG_OK;
}

extern "C" double mg_nested_log_total_error(double log_ev, double log_dev, int32_t nlive) {
  const double log_rel_error2 = -std::log((double)nlive);          // nested.ml:148-150
  return 0.5 * h_log_sum_logs(2.0 * log_dev, log_rel_error2 + 2.0 * log_ev);
}

extern "C" int mg_nested_weights(mg_ctx *ctx, const double *ll, int64_t n, int32_t nlive, int32_t batch,
                                 double *log_ev, double *log_dev, double *logw) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, ll && log_ev && log_dev && logw, "nested_weights: null argument");
  MG_REQUIRE(ctx, nlive >= 2 && n >= nlive && batch >= 1 && batch < nlive, "nested_weights: bad sizes");
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf<double> d_ll, d_w;
  MG_CUDA(ctx, upload(d_ll, ll, (size_t)n, ctx->stream));
  MG_CUDA(ctx, d_w.alloc(n, ctx->stream));
  int rc = weights_dev(ctx, d_ll.get(), n, nlive, batch, 0.0, log_ev, log_dev, d_w.get());
  if (rc) return rc;
  MG_CUDA(ctx, cudaMemcpyAsync(logw, d_w.get(), sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MG_OK;
}

extern "C" int mg_nested_evidence(mg_ctx *ctx, const mg_logfn *like, const mg_logfn *prior, const double *prior_lo,
                                  const double *prior_hi, const mg_nested_cfg *cfg, double *log_ev, double *log_dev,
                                  int64_t *npts, double *pts, double *ll, double *lp, double *logw) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, like && prior && prior_lo && prior_hi && cfg && log_ev && log_dev && npts && ll && logw,
             "nested_evidence: null argument");
  const int D = cfg->dim, nlive = cfg->nlive, K = cfg->batch;
  MG_REQUIRE(ctx, D >= 1 && D <= 64, "nested_evidence: dim must be in 1..64");
  MG_REQUIRE(ctx, nlive >= 2 && cfg->nmcmc >= 0, "nested_evidence: bad nlive / nmcmc");
  MG_REQUIRE(ctx, K >= 1 && K < nlive, "nested_evidence: need 1 <= batch < nlive");
  MG_REQUIRE(ctx, cfg->max_points >= nlive + K, "nested_evidence: max_points too small");
  int rc;
  if ((rc = validate_logfn(ctx, like, D, "log_likelihood"))) return rc;
  if ((rc = validate_logfn(ctx, prior, D, "log_prior"))) return rc;
  MG_REQUIRE(ctx, like->kind < MG_FN_USER && prior->kind < MG_FN_USER,
             "nested_evidence: run-time plugins are supported by mcmc_array and logfn_eval only");
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevLogFn dl, dp;
  MG_CUDA(ctx, dl.upload_from(like, s));
  MG_CUDA(ctx, dp.upload_from(prior, s));
  DevBuf<double> d_plo, d_phi, lx[2], lll[2], llp[2], fx, fll, flp, rx, rll, rlp;
  DevBuf<uint64_t> keys;
  DevBuf<int32_t> order;
  DevBuf<int> d_fail;
  MG_CUDA(ctx, upload(d_plo, prior_lo, D, s));
  MG_CUDA(ctx, upload(d_phi, prior_hi, D, s));
  for (int b = 0; b < 2; ++b) {
    MG_CUDA(ctx, lx[b].alloc((size_t)nlive * D, s));
    MG_CUDA(ctx, lll[b].alloc(nlive, s));
    MG_CUDA(ctx, llp[b].alloc(nlive, s));
  }
  MG_CUDA(ctx, fx.alloc((size_t)K * D, s)); MG_CUDA(ctx, fll.alloc(K, s)); MG_CUDA(ctx, flp.alloc(K, s));
  const int64_t cap = cfg->max_points;
  MG_CUDA(ctx, rx.alloc((size_t)cap * D, s)); MG_CUDA(ctx, rll.alloc(cap, s)); MG_CUDA(ctx, rlp.alloc(cap, s));
  MG_CUDA(ctx, keys.alloc(nlive, s)); MG_CUDA(ctx, order.alloc(nlive, s));
  MG_CUDA(ctx, d_fail.alloc(1, s));
  MG_CUDA(ctx, cudaMemsetAsync(d_fail.get(), 0, sizeof(int), s));

  NestArgs a{};
  a.like = dl.params; a.prior = dp.params; a.plo = d_plo.get(); a.phi = d_phi.get();
  a.key = next_key(ctx);
  a.D = D; a.nlive = nlive; a.K = K; a.nmcmc = cfg->nmcmc;
  a.mode_hop = cfg->mode_hopping_frac;
  a.de_sigma = 2.38 / std::sqrt(2.0 * (double)D);   // mcmc.ml:212
  a.fresh_x = fx.get(); a.fresh_ll = fll.get(); a.fresh_lp = flp.get(); a.fail = d_fail.get();

#define MG_NEST_DISPATCH(KERNEL, GRID, BLOCK, ...)                                   \
  do {                                                                               \
    if (D <= 2) KERNEL<2><<<GRID, BLOCK, 0, s>>>(__VA_ARGS__);                       \
    else if (D <= 4) KERNEL<4><<<GRID, BLOCK, 0, s>>>(__VA_ARGS__);                  \
    else if (D <= 8) KERNEL<8><<<GRID, BLOCK, 0, s>>>(__VA_ARGS__);                  \
    else if (D <= 16) KERNEL<16><<<GRID, BLOCK, 0, s>>>(__VA_ARGS__);                \
    else if (D <= 32) KERNEL<32><<<GRID, BLOCK, 0, s>>>(__VA_ARGS__);                \
    else KERNEL<64><<<GRID, BLOCK, 0, s>>>(__VA_ARGS__);                             \
  } while (0)

  int cur = 0;
  MG_NEST_DISPATCH(nest_init_kernel, (nlive + 127) / 128, 128, a, lx[1].get(), lll[1].get(), llp[1].get());
  MG_CHECK_LAUNCH(ctx);
  // Array.fast_sort by log_likelihood (:132): stable radix sort + gather
  auto sort_live = [&](int from, int to) -> int {
    nest_keys_kernel<<<(nlive + 255) / 256, 256, 0, s>>>(lll[from].get(), nlive, keys.get(), order.get());
    MG_CHECK_LAUNCH(ctx);
    int r = radix_sort_pairs(ctx, keys.get(), order.get(), nlive, 1);
    if (r) return r;
    nest_gather_kernel<<<(nlive + 255) / 256, 256, 0, s>>>(lx[from].get(), lll[from].get(), llp[from].get(), order.get(),
                                                          nlive, D, lx[to].get(), lll[to].get(), llp[to].get());
    MG_CHECK_LAUNCH(ctx);
    return MG_OK;
  };
  if ((rc = sort_live(1, 0))) return rc;
  cur = 0;

  Shrink sh(nlive, K);
  double log_vol = 0.0, log_int = -INFINITY;
  int64_t R = 0;
  std::vector<double> h_low(K);
  double h_edge[2];
  const int rblock = 64;
  time_begin(ctx);
  for (;;) {
    if (R + K + nlive > cap) return set_err(ctx, MG_EFAIL, "nested_evidence: max_points too small");
    // threshold = ll of the K-th lowest live point; the K lowest are retired
    MG_CUDA(ctx, cudaMemcpyAsync(h_low.data(), lll[cur].get(), sizeof(double) * K, cudaMemcpyDeviceToHost, s));
    MG_CUDA(ctx, cudaStreamSynchronize(s));
    a.live_x = lx[cur].get(); a.live_ll = lll[cur].get(); a.live_lp = llp[cur].get();
    a.R = R; a.threshold = h_low[K - 1];
    MG_NEST_DISPATCH(nest_replace_kernel, (K + rblock - 1) / rblock, rblock, a);
    MG_CHECK_LAUNCH(ctx);
    // retired_pt :: retired_pts (:137)
    MG_CUDA(ctx, cudaMemcpyAsync(rx.get() + R * D, lx[cur].get(), sizeof(double) * K * D, cudaMemcpyDeviceToDevice, s));
    MG_CUDA(ctx, cudaMemcpyAsync(rll.get() + R, lll[cur].get(), sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
    MG_CUDA(ctx, cudaMemcpyAsync(rlp.get() + R, llp[cur].get(), sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
    // replace_live_point (:26-43): new points take the vacated front slots, then a stable sort
    MG_CUDA(ctx, cudaMemcpyAsync(lx[cur].get(), fx.get(), sizeof(double) * K * D, cudaMemcpyDeviceToDevice, s));
    MG_CUDA(ctx, cudaMemcpyAsync(lll[cur].get(), fll.get(), sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
    MG_CUDA(ctx, cudaMemcpyAsync(llp[cur].get(), flp.get(), sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
    if ((rc = sort_live(cur, 1 - cur))) return rc;
    cur = 1 - cur;
    // running tracker (:138-141), quirk F5d kept: log_dv = log_vol +. vol_fraction
    for (int j = 0; j < K; ++j) {
      const double vol_fraction = 1.0 / (double)(nlive - j);
      const double log_new_vol = log_vol + std::log1p(-vol_fraction);
      const double log_dv = log_vol + vol_fraction;
      log_int = h_log_sum_logs(log_int, h_low[j] + log_dv);
      log_vol = log_new_vol;
    }
    R += K;
    int h_fail = 0;
    MG_CUDA(ctx, cudaMemcpyAsync(h_edge, lll[cur].get() + (nlive - 1), sizeof(double), cudaMemcpyDeviceToHost, s));
    MG_CUDA(ctx, cudaMemcpyAsync(&h_fail, d_fail.get(), sizeof(int), cudaMemcpyDeviceToHost, s));
    MG_CUDA(ctx, cudaStreamSynchronize(s));
    if (h_fail) return set_err(ctx, MG_EFAIL, "Error in draw_new_live_point: new log(L) below the threshold");  // :70-72
    // remaining_integral_negligable (:45-48)
    const double log_live_estimate = log_vol + h_edge[0];
    if (log_live_estimate - h_log_sum_logs(log_int, log_live_estimate) <= std::log(cfg->epsrel)) break;
  }
  time_end(ctx);
  const int64_t n = R + nlive;
  if (n > cap) return set_err(ctx, MG_EFAIL, "nested_evidence: max_points too small");
  // all points ascending in ll: retired (in order) followed by the sorted live set (:143)
  MG_CUDA(ctx, cudaMemcpyAsync(rx.get() + R * D, lx[cur].get(), sizeof(double) * nlive * D, cudaMemcpyDeviceToDevice, s));
  MG_CUDA(ctx, cudaMemcpyAsync(rll.get() + R, lll[cur].get(), sizeof(double) * nlive, cudaMemcpyDeviceToDevice, s));
  MG_CUDA(ctx, cudaMemcpyAsync(rlp.get() + R, llp[cur].get(), sizeof(double) * nlive, cudaMemcpyDeviceToDevice, s));
  DevBuf<double> d_w;
  MG_CUDA(ctx, d_w.alloc(n, s));
  if ((rc = weights_dev(ctx, rll.get(), n, nlive, K, 0.0, log_ev, log_dev, d_w.get()))) return rc;
  if (pts) MG_CUDA(ctx, cudaMemcpyAsync(pts, rx.get(), sizeof(double) * n * D, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaMemcpyAsync(ll, rll.get(), sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (lp) MG_CUDA(ctx, cudaMemcpyAsync(lp, rlp.get(), sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaMemcpyAsync(logw, d_w.get(), sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  *npts = n;
  return MG_OK;
}
