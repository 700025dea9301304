// og_models.hpp -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
//
// CPU restatement of stats.ml and of the likelihood / prior / proposal
// definitions the reference ships in bin/ and test/ (the "built-in plugins" of
// include/mcmc_gpu.h).  Every function cites the reference lines it follows.
// Arithmetic is written operation by operation in the reference's order;
// compile with -ffp-contract=off so the compiler adds no fused multiply-adds
// (ocamlopt emits none on x86-64).  The only explicit fma()s are in
// MG_FN_GAUSS_CORR, a model this project defines (BASELINE.json config 2),
// and in the MG_PROP_BOX proposal (see there).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

#include "../include/mcmc_gpu.h"
#include "og_rng.hpp"

namespace og {

static const double NEG_INF = -std::numeric_limits<double>::infinity();
static const double PI = 4.0 * std::atan(1.0);  // stats.ml:56

// ---- stats.ml -------------------------------------------------------------

// stats.ml:17-23
inline double mean(const double *xs, int64_t n) {
  double sum = 0.0;
  for (int64_t i = 0; i < n; ++i) sum = sum + xs[i];
  return sum / (double)n;
}
// stats.ml:35-43
inline double stdev(const double *xs, int64_t n, bool have_mean, double mu) {
  double m = have_mean ? mu : mean(xs, n);
  double var = 0.0;
  for (int64_t i = 0; i < n; ++i) { double x = xs[i] - m; var = var + x * x; }
  return std::sqrt(var / (double)(n - 1));
}
// stats.ml:58-70   xs: [n][D]
inline void multi_mean(const double *xs, int64_t n, int D, double *mu) {
  for (int j = 0; j < D; ++j) mu[j] = 0.0;
  for (int64_t i = 0; i < n; ++i)
    for (int j = 0; j < D; ++j) mu[j] = mu[j] + xs[i * D + j];
  for (int j = 0; j < D; ++j) mu[j] = mu[j] / (double)n;
}
// stats.ml:72-87
inline void multi_std(const double *xs, int64_t n, int D, const double *mean_or_null,
                      double *sd) {
  std::vector<double> mu(D);
  if (mean_or_null) for (int j = 0; j < D; ++j) mu[j] = mean_or_null[j];
  else multi_mean(xs, n, D, mu.data());
  for (int j = 0; j < D; ++j) sd[j] = 0.0;
  for (int64_t i = 0; i < n; ++i)
    for (int j = 0; j < D; ++j) { double dx = xs[i * D + j] - mu[j]; sd[j] = sd[j] + dx * dx; }
  for (int j = 0; j < D; ++j) sd[j] = std::sqrt(sd[j] / (double)(n - 1));
}
// stats.ml:93-96
inline double log_cauchy(double x0, double gamma, double x) {
  double dx = (x - x0) / gamma;
  return 0.0 - std::log(PI * gamma) - std::log(1.0 + dx * dx);
}
// stats.ml:98-101
inline double log_gaussian(double mu, double sigma, double x) {
  double dx = (x - mu) / sigma;
  return -0.91893853320467274178 - std::log(sigma) - 0.5 * dx * dx;
}
// stats.ml:103-108
inline double log_multi_gaussian(const double *mu, const double *sigma, const double *x, int D) {
  double result = 0.0;
  for (int i = 0; i < D; ++i) result = result + log_gaussian(mu[i], sigma[i], x[i]);
  return result + 0.0;
}
// stats.ml:113-124 (Leva's ratio of uniforms)
inline double draw_gaussian(Rng &r, double mu, double sigma) {
  for (;;) {
    double u = r.uniform();
    double v = 1.7156 * (r.uniform() - 0.5);
    double x = u - 0.449871;
    double y = std::fabs(v) + 0.386595;
    double q = x * x + y * (0.19600 * y - 0.25472 * x);
    if (q > 0.27597 && (q > 0.27846 || v * v > (-4.0) * std::log(u) * u * u)) continue;
    return mu + sigma * v / u;
  }
}
// stats.ml:89-91
inline double draw_cauchy(Rng &r, double x0, double gamma) {
  double p = r.uniform();
  return x0 + gamma * std::tan(PI * (p - 0.5));
}
// stats.ml:126-128
inline double draw_uniform(Rng &r, double a, double b) {
  double d = b - a;
  return a + d * r.uniform();
}
// stats.ml:215-221
inline double log_lognormal(double mu, double sigma, double x) {
  double lx = std::log(x);
  double d = (lx - mu) / sigma;
  double ls = std::log(sigma);
  return -0.91893853320467274178 - lx - ls - 0.5 * d * d;
}
// stats.ml:240-248
inline double log_sum_logs(double a, double b) {
  if (a == NEG_INF && b == NEG_INF) return NEG_INF;
  if (b > a) { double t = a; a = b; b = t; }
  double r = std::exp(b - a);
  return a + std::log1p(r);
}
// mcmc.ml:155-163 (the Mcmc copy uses log (1 + exp) instead of log1p)
inline double mcmc_log_sum_logs(double la, double lb) {
  if (la == NEG_INF && lb == NEG_INF) return NEG_INF;
  if (la > lb) { double lr = lb - la; return la + std::log(1.0 + std::exp(lr)); }
  double lr = la - lb; return lb + std::log(1.0 + std::exp(lr));
}
// stats.ml:223-238 for lags 0..nslides-1 (the reference loop runs
// one lag too far and writes out of bounds; parity unpinned).
inline void slow_autocorrelation(const double *x, int64_t n, int nslides, double *result) {
  double mu = mean(x, n);
  double sigma = stdev(x, n, false, 0.0);
  double sigma2 = sigma * sigma;
  for (int i = 0; i < nslides; ++i) {
    double acc = 0.0;
    for (int64_t j = 0; j <= n - 1 - i; ++j) {
      double dx = x[j] - mu, dxshift = x[j + i] - mu;
      acc = acc + dx * dxshift / sigma2;
    }
    result[i] = acc / (double)(n - i);
  }
}

// ---- built-in log-density plugins -----------------------------------------

struct LogFn {
  int kind = MG_FN_ZERO, D = 0;
  double scale = 1.0;
  std::vector<double> p;
  LogFn() {}
  explicit LogFn(const mg_logfn *f) : kind(f->kind), D(f->dim), scale(f->scale) {
    if (f->nparams > 0) p.assign(f->params, f->params + f->nparams);
    int64_t need = -1;
    switch (kind) {
      case MG_FN_ZERO: need = 0; break;
      case MG_FN_CONST: need = 1; break;
      case MG_FN_BOX_CLOSED: case MG_FN_BOX_OPEN: need = 2 * D + 1; break;
      case MG_FN_GAUSS_DIAG: need = 2 * D; break;
      case MG_FN_GAUSS_CORR: need = D + (int64_t)D * (D + 1) / 2 + 1; break;
      case MG_FN_GAUSS_DATA: case MG_FN_CAUCHY_DATA:
        if (D != 2) throw std::invalid_argument("data likelihoods need dim 2");
        need = -1; break;
      case MG_FN_SHELL: need = D + 2; break;
      case MG_FN_GAUSS_MIX:
        if (p.empty()) throw std::invalid_argument("mix: no params");
        need = 1 + (int64_t)p[0] * D + D; break;
      default: throw std::invalid_argument("unknown log-density kind");
    }
    if (need >= 0 && (int64_t)p.size() != need) throw std::invalid_argument("bad nparams");
  }
  double raw(const double *x) const {
    switch (kind) {
      case MG_FN_ZERO: return 0.0;
      case MG_FN_CONST: return p[0];
      case MG_FN_BOX_CLOSED: {  // bin/gaussian_cauchy_efficiency.ml:60-67
        for (int i = 0; i < D; ++i) if (x[i] < p[i] || x[i] > p[D + i]) return NEG_INF;
        return p[2 * D];
      }
      case MG_FN_BOX_OPEN: {  // test/nested_test.ml:24-28
        for (int i = 0; i < D; ++i) if (!(x[i] > p[i] && x[i] < p[D + i])) return NEG_INF;
        return p[2 * D];
      }
      case MG_FN_GAUSS_DIAG: return log_multi_gaussian(&p[0], &p[D], x, D);
      case MG_FN_GAUSS_CORR: {
        const double *mu = &p[0], *L = &p[D];
        double q = 0.0; int k = 0;
        for (int i = 0; i < D; ++i) {
          double y = L[k] * (x[0] - mu[0]); ++k;
          for (int j = 1; j <= i; ++j, ++k) y = std::fma(L[k], x[j] - mu[j], y);
          q = std::fma(y, y, q);
        }
        return std::fma(-0.5, q, p[D + D * (D + 1) / 2]);
      }
      case MG_FN_GAUSS_DATA: {  // bin/gaussian_cauchy_efficiency.ml:69-77
        double sum = 0.0;
        for (size_t i = 0; i < p.size(); ++i) sum = sum + log_gaussian(x[0], x[1], p[i]);
        return sum + 0.0;
      }
      case MG_FN_CAUCHY_DATA: {  // bin/gaussian_cauchy_efficiency.ml:79-87
        double sum = 0.0;
        for (size_t i = 0; i < p.size(); ++i) sum = sum + log_cauchy(x[0], x[1], p[i]);
        return sum + 0.0;
      }
      case MG_FN_SHELL: {
        double s = 0.0;
        for (int i = 0; i < D; ++i) { double dx = x[i] - p[i]; s = s + dx * dx; }
        return log_gaussian(p[D], p[D + 1], std::sqrt(s));
      }
      case MG_FN_GAUSS_MIX: {  // test/nested_test.ml:47-53
        int K = (int)p[0]; double tot = 0.0;
        for (int k = 0; k < K; ++k)
          tot = tot + std::exp(log_multi_gaussian(&p[1 + k * D], &p[1 + K * D], x, D));
        return std::log(tot);
      }
    }
    return NEG_INF;
  }
  double operator()(const double *x) const {
    double v = raw(x);
    return scale == 1.0 ? v : scale * v;
  }
};

// ---- built-in jump proposals ----------------------------------------------

struct Proposal {
  int kind = MG_PROP_BOX, D = 0;
  std::vector<double> p;
  std::vector<double> w;          // MG_PROP_MIXTURE: normalised weights
  std::vector<Proposal> comps;    // MG_PROP_MIXTURE: components
  Proposal() {}
  Proposal(int kind_, int D_, const double *pp, int64_t np) : kind(kind_), D(D_) { init(pp, np); }
  explicit Proposal(const mg_proposal *f) : kind(f->kind), D(f->dim) { init(f->params, f->nparams); }
  void init(const double *pp, int64_t np) {
    if (np > 0) p.assign(pp, pp + np);
    int64_t need = 0;
    switch (kind) {
      case MG_PROP_BOX: need = D; break;
      case MG_PROP_WRAP: need = 3 * D; break;
      case MG_PROP_INDEP_GAUSS: need = 2 * D; break;
      case MG_PROP_LEFT_BIASED: need = 1; if (D != 1) throw std::invalid_argument("left-biased is 1-D"); break;
      case MG_PROP_ONE_SIDED: need = 2; if (D != 1) throw std::invalid_argument("one-sided is 1-D"); break;
      case MG_PROP_MIXTURE: {  // combine_jump_proposals mcmc.ml:165-167
        if (p.empty()) throw std::invalid_argument("mixture: no params");
        int K = (int)p[0]; size_t k = 1; double ptot = 0.0;
        for (int c = 0; c < K; ++c) {
          if (k + 3 > p.size()) throw std::invalid_argument("mixture: truncated params");
          double wt = p[k]; int ck = (int)p[k + 1]; int64_t cn = (int64_t)p[k + 2];
          if (ck == MG_PROP_MIXTURE || k + 3 + cn > p.size()) throw std::invalid_argument("mixture: bad component");
          ptot = ptot + wt; w.push_back(wt);
          comps.emplace_back(ck, D, p.data() + k + 3, cn);
          k += 3 + cn;
        }
        for (auto &x : w) x = x / ptot;
        need = (int64_t)k; break;
      }
      case MG_PROP_DE: {       // mode_hopping_frac, M, samples[M][D]
        if (p.size() < 2 || !(p[1] >= 2.0)) throw std::invalid_argument("differential_evolution_proposal: need at least two samples");
        need = 2 + (int64_t)p[1] * D; break;
      }
      default: throw std::invalid_argument("unknown proposal kind");
    }
    if ((int64_t)p.size() != need) throw std::invalid_argument("bad proposal nparams");
  }
  // mcmc.ml:187-196 (reflects)
  static double uniform_wrapping(Rng &r, double xmin, double xmax, double dx, double x) {
    double delta_x = (r.uniform() - 0.5) * dx;
    double new_x = x + delta_x;
    for (;;) {
      if (new_x < xmin) new_x = xmin + (xmin - new_x);
      else if (new_x >= xmax) new_x = xmax - (new_x - xmax);
      else return new_x;
    }
  }
  void propose(Rng &r, const double *x, double *y) const {
    switch (kind) {
      case MG_PROP_BOX:  // bin/evidence_direct.ml:24-25,39-43: x + random_between (-h) h.
        // The reference evaluates a +. (b -. a) *. u with two roundings.  This plugin (a
        // model-level definition, not library code) uses the raw mantissa draw
        // m = 1 + u in [1,2) and ONE fused multiply-add: a + (b-a) u = (a - (b-a)) + (b-a) m.
        // Same distribution to within an ulp per draw; 21 fewer FP64 instructions per
        // 10-D step on the GPU, whose plugin does exactly the same arithmetic.
        for (int i = 0; i < D; ++i) {
          double a = -p[i], b = p[i];
          double w = b - a, c = a - w;
          y[i] = x[i] + std::fma(w, r.uniform12(), c);
        }
        break;
      case MG_PROP_WRAP:
        for (int i = 0; i < D; ++i) y[i] = uniform_wrapping(r, p[i], p[D + i], p[2 * D + i], x[i]);
        break;
      case MG_PROP_INDEP_GAUSS:  // test/mcmc_test.ml:119-122
        for (int i = 0; i < D; ++i) y[i] = draw_gaussian(r, p[i], p[D + i]);
        break;
      case MG_PROP_LEFT_BIASED:  // test/mcmc_test.ml:66-70
        if (r.uniform() < 0.75) y[0] = x[0] - p[0] * r.uniform();
        else y[0] = x[0] + p[0] * r.uniform();
        break;
      case MG_PROP_ONE_SIDED:  // test/mcmc_test.ml:186-189
        y[0] = x[0] + p[0] * (p[1] * r.uniform());
        break;
      case MG_PROP_DE: {       // differential_evolution_proposal, mcmc.ml:198-218
        const double mode_hop = p[0];
        const uint64_t M = (uint64_t)p[1];
        uint64_t i = r.below(M), j;                              // :201-202
        do { j = r.below(M); } while (j == i);
        double d;
        if (mode_hop != 0.0 && r.uniform() < mode_hop) d = 1.0;  // :209-210
        else { double sigma = 2.38 / std::sqrt(2.0 * (double)D); d = draw_gaussian(r, 0.0, sigma); }   // :212-213
        const double *xs = p.data() + 2 + i * (uint64_t)D, *ys = p.data() + 2 + j * (uint64_t)D;
        for (int k = 0; k < D; ++k) y[k] = x[k] + d * (ys[k] - xs[k]);   // :215-217
        break;
      }
      case MG_PROP_MIXTURE: {  // mcmc.ml:168-176
        double prob = r.uniform();
        size_t c = 0;
        for (; c < comps.size(); ++c) { if (prob < w[c]) break; prob = prob - w[c]; }
        if (c == comps.size()) c = comps.size() - 1;  // the reference raises Failure here (prob ~ 1e-16)
        comps[c].propose(r, x, y);
        break;
      }
    }
  }
  // log q(x -> y)
  double log_q(const double *x, const double *y) const {
    switch (kind) {
      case MG_PROP_INDEP_GAUSS: {  // test/mcmc_test.ml:123-126
        double s = 0.0;
        for (int i = 0; i < D; ++i) s = s + log_gaussian(p[i], p[D + i], y[i]);
        return s;
      }
      case MG_PROP_LEFT_BIASED:  // test/mcmc_test.ml:73
        return x[0] > y[0] ? std::log(0.75) : std::log(0.25);
      case MG_PROP_ONE_SIDED: {  // test/mcmc_test.ml:190-199
        double d = p[0] * (y[0] - x[0]);
        return (d >= 0.0 && d <= p[1]) ? 0.0 - std::log(p[1]) : NEG_INF;
      }
      case MG_PROP_MIXTURE: {  // mcmc.ml:177-184
        double log_jump = NEG_INF;
        for (size_t c = 0; c < comps.size(); ++c) {
          double log_local = std::log(w[c]) + comps[c].log_q(x, y);
          log_jump = mcmc_log_sum_logs(log_jump, log_local);
        }
        return log_jump;
      }
      default: return 0.0;
    }
  }
};

}  // namespace og
