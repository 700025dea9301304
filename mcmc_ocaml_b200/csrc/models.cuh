// models.cuh -- device-side log-density and jump-proposal plugins.
//
// The reference takes OCaml closures for log_likelihood / log_prior /
// jump_proposal / log_jump_prob (mcmc.mli:58-60).  On the GPU these are
// plugin *types* compiled into the sampler kernels (no indirect calls in the
// hot loop):
//   - static plugins (GaussCorr, ZeroFn, BoxProp): parameters travel in the
//     kernel argument block (constant bank) and the dimension is a template
//     parameter, for the configurations whose throughput is measured;
//   - dynamic plugins (DynFn, DynProp): one type that switches on the
//     registered kind id, parameters in global memory, for everything else.
// Arithmetic follows stats.ml / the bin/ and test/ model definitions
// operation by operation (file is compiled with -fmad=false; the only fused
// multiply-adds are the explicit fma() calls of the GAUSS_CORR model).
#pragma once
#include <cstdint>

#include "../../include/mcmc_gpu.h"
#include "rng.cuh"

namespace mg {

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xFFF0000000000000ll); }
__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7FF8000000000000ll); }

// Dynamic plugins unroll their per-dimension loops only for small DMAX (state
// in registers); larger DMAX keep rolled loops over local-memory arrays.
#define MG_DYN_UNROLL(DMAX) ((DMAX) <= 8 ? (DMAX) : 1)

#define MG_PI 3.14159265358979311600  /* 4.0 *. atan 1.0, stats.ml:56 */

// stats.ml:98-101
__device__ __forceinline__ double log_gaussian(double mu, double sigma, double x) {
  const double dx = (x - mu) / sigma;
  return -0.91893853320467274178 - log(sigma) - 0.5 * dx * dx;
}
// the same with log sigma already evaluated (it is loop invariant in the
// data likelihoods, bin/gaussian_cauchy_efficiency.ml:69-77)
__device__ __forceinline__ double log_gaussian_ls(double mu, double sigma, double log_sigma, double x) {
  const double dx = (x - mu) / sigma;
  return -0.91893853320467274178 - log_sigma - 0.5 * dx * dx;
}
// stats.ml:93-96
__device__ __forceinline__ double log_cauchy_lg(double x0, double gamma, double log_pi_gamma, double x) {
  const double dx = (x - x0) / gamma;
  return 0.0 - log_pi_gamma - log(1.0 + dx * dx);
}
// stats.ml:240-248
__device__ __forceinline__ double log_sum_logs(double a, double b) {
  if (a == neg_inf() && b == neg_inf()) return neg_inf();
  if (b > a) { const double t = a; a = b; b = t; }
  const double r = exp(b - a);
  return a + log1p(r);
}
// stats.ml:113-124 (Leva)
__device__ __forceinline__ double draw_gaussian(Rng &r, double mu, double sigma) {
  for (;;) {
    const double u = r.uniform();
    const double v = 1.7156 * (r.uniform() - 0.5);
    const double x = u - 0.449871;
    const double y = fabs(v) + 0.386595;
    const double q = x * x + y * (0.19600 * y - 0.25472 * x);
    if (q > 0.27597 && (q > 0.27846 || v * v > (-4.0) * log(u) * u * u)) continue;
    return mu + sigma * v / u;
  }
}
// stats.ml:126-128
__device__ __forceinline__ double draw_uniform(Rng &r, double a, double b) {
  const double d = b - a;
  return a + d * r.uniform();
}
// mcmc.ml:187-196 (reflects at the bounds)
__device__ __forceinline__ double uniform_wrapping(Rng &r, double xmin, double xmax, double dx, double x) {
  const double delta_x = (r.uniform() - 0.5) * dx;
  double new_x = x + delta_x;
  for (;;) {
    if (new_x < xmin) new_x = xmin + (xmin - new_x);
    else if (new_x >= xmax) new_x = xmax - (new_x - xmax);
    else return new_x;
  }
}

// ---------------------------------------------------------------------------
// static plugins
// ---------------------------------------------------------------------------

struct ZeroFn {
  struct Params {};
  template <int D>
  static __device__ __forceinline__ double eval(const Params &, const double (&)[D], int) { return 0.0; }
};

// MG_FN_GAUSS_CORR: logc - 1/2 |L (x - mu)|^2, L lower triangular, packed by rows.
template <int D>
struct GaussCorr {
  struct Params { double mu[D]; double L[D * (D + 1) / 2]; double logc; };
  template <int DD>
  static __device__ __forceinline__ double eval(const Params &p, const double (&x)[DD], int) {
    static_assert(DD == D, "GaussCorr: dimension mismatch");
    double z[D];
#pragma unroll
    for (int j = 0; j < D; ++j) z[j] = x[j] - p.mu[j];
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double y = p.L[i * (i + 1) / 2] * z[0];
#pragma unroll
      for (int j = 1; j <= i; ++j) y = fma(p.L[i * (i + 1) / 2 + j], z[j], y);
      q = fma(y, y, q);
    }
    return fma(-0.5, q, p.logc);
  }
};

// MG_PROP_BOX: x_i + random_between (-h_i) h_i  (bin/evidence_direct.ml:24-43)
template <int D>
struct BoxProp {
  struct Params { double h[D]; };
  static constexpr bool kSymmetric = true;
  template <int DD>
  static __device__ __forceinline__ void propose(const Params &p, Rng &r, const double (&x)[DD],
                                                 double (&y)[DD], int) {
    static_assert(DD == D, "BoxProp: dimension mismatch");
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double a = -p.h[i], b = p.h[i];
      y[i] = x[i] + (a + (b - a) * r.uniform());
    }
  }
  template <int DD>
  static __device__ __forceinline__ double log_q(const Params &, const double (&)[DD], const double (&)[DD], int) {
    return 0.0;
  }
};

// ---------------------------------------------------------------------------
// dynamic plugins (switch on the registered kind; params in global memory)
// ---------------------------------------------------------------------------

struct DynFnParams {
  int32_t kind, dim;
  double scale;
  const double *p;  // device
  int64_t np;
};

template <int DMAX>
__device__ __forceinline__ double dyn_log_multi_gaussian(const double *mu, const double *sigma,
                                                         const double (&x)[DMAX], int d) {
  double result = 0.0;  // stats.ml:103-108
#pragma unroll (DMAX <= 8 ? DMAX : 1)
  for (int i = 0; i < DMAX; ++i)
    if (i < d) result = result + log_gaussian(__ldg(mu + i), __ldg(sigma + i), x[i]);
  return result + 0.0;
}

struct DynFn {
  typedef DynFnParams Params;
  template <int DMAX>
  static __device__ __forceinline__ double raw(const Params &f, const double (&x)[DMAX], int d) {
    const double *p = f.p;
    switch (f.kind) {
      case MG_FN_ZERO: return 0.0;
      case MG_FN_CONST: return __ldg(p);
      case MG_FN_BOX_CLOSED: {  // bin/gaussian_cauchy_efficiency.ml:60-67
        bool out = false;
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) out = out || (x[i] < __ldg(p + i)) || (x[i] > __ldg(p + d + i));
        return out ? neg_inf() : __ldg(p + 2 * d);
      }
      case MG_FN_BOX_OPEN: {  // test/nested_test.ml:24-28
        bool in = true;
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) in = in && (x[i] > __ldg(p + i)) && (x[i] < __ldg(p + d + i));
        return in ? __ldg(p + 2 * d) : neg_inf();
      }
      case MG_FN_GAUSS_DIAG: return dyn_log_multi_gaussian<DMAX>(p, p + d, x, d);
      case MG_FN_GAUSS_CORR: {
        const double *mu = p, *L = p + d;
        double z[DMAX];
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int j = 0; j < DMAX; ++j) z[j] = (j < d) ? x[j] - __ldg(mu + j) : 0.0;
        double q = 0.0;
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i) {
          if (i < d) {
            double y = __ldg(L + i * (i + 1) / 2) * z[0];
#pragma unroll (DMAX <= 8 ? DMAX : 1)
            for (int j = 1; j < DMAX; ++j)
              if (j <= i) y = fma(__ldg(L + i * (i + 1) / 2 + j), z[j], y);
            q = fma(y, y, q);
          }
        }
        return fma(-0.5, q, __ldg(p + d + d * (d + 1) / 2));
      }
      case MG_FN_GAUSS_DATA: {  // bin/gaussian_cauchy_efficiency.ml:69-77
        const double mu = x[0], sigma = x[1], ls = log(sigma);
        double sum = 0.0;
        for (int64_t i = 0; i < f.np; ++i) sum = sum + log_gaussian_ls(mu, sigma, ls, __ldg(p + i));
        return sum + 0.0;
      }
      case MG_FN_CAUCHY_DATA: {  // bin/gaussian_cauchy_efficiency.ml:79-87
        const double x0 = x[0], gamma = x[1], lg = log(MG_PI * gamma);
        double sum = 0.0;
        for (int64_t i = 0; i < f.np; ++i) sum = sum + log_cauchy_lg(x0, gamma, lg, __ldg(p + i));
        return sum + 0.0;
      }
      case MG_FN_SHELL: {
        double s = 0.0;
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) { const double dx = x[i] - __ldg(p + i); s = s + dx * dx; }
        return log_gaussian(__ldg(p + d), __ldg(p + d + 1), sqrt(s));
      }
      case MG_FN_GAUSS_MIX: {  // test/nested_test.ml:47-53
        const int K = (int)__ldg(p);
        double tot = 0.0;
        for (int k = 0; k < K; ++k) tot = tot + exp(dyn_log_multi_gaussian<DMAX>(p + 1 + k * d, p + 1 + K * d, x, d));
        return log(tot);
      }
    }
    return neg_inf();
  }
  template <int DMAX>
  static __device__ __forceinline__ double eval(const Params &f, const double (&x)[DMAX], int d) {
    const double v = raw<DMAX>(f, x, d);
    return f.scale == 1.0 ? v : f.scale * v;
  }
};

struct DynPropParams {
  int32_t kind, dim;
  const double *p;  // device
  int64_t np;
};

struct DynProp {
  typedef DynPropParams Params;
  static constexpr bool kSymmetric = false;
  template <int DMAX>
  static __device__ __forceinline__ void propose(const Params &f, Rng &r, const double (&x)[DMAX],
                                                 double (&y)[DMAX], int d) {
    const double *p = f.p;
    switch (f.kind) {
      case MG_PROP_BOX:
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) { const double a = -__ldg(p + i), b = __ldg(p + i); y[i] = x[i] + (a + (b - a) * r.uniform()); }
        break;
      case MG_PROP_WRAP:
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) y[i] = uniform_wrapping(r, __ldg(p + i), __ldg(p + d + i), __ldg(p + 2 * d + i), x[i]);
        break;
      case MG_PROP_INDEP_GAUSS:
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) y[i] = draw_gaussian(r, __ldg(p + i), __ldg(p + d + i));
        break;
      case MG_PROP_LEFT_BIASED:  // test/mcmc_test.ml:66-70
        if (r.uniform() < 0.75) y[0] = x[0] - __ldg(p) * r.uniform();
        else y[0] = x[0] + __ldg(p) * r.uniform();
        break;
    }
#pragma unroll (DMAX <= 8 ? DMAX : 1)
    for (int i = 0; i < DMAX; ++i)
      if (i >= d) y[i] = 0.0;
  }
  template <int DMAX>
  static __device__ __forceinline__ double log_q(const Params &f, const double (&x)[DMAX],
                                                 const double (&y)[DMAX], int d) {
    switch (f.kind) {
      case MG_PROP_INDEP_GAUSS: {  // test/mcmc_test.ml:123-126
        double s = 0.0;
#pragma unroll (DMAX <= 8 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) s = s + log_gaussian(__ldg(f.p + i), __ldg(f.p + d + i), y[i]);
        return s;
      }
      case MG_PROP_LEFT_BIASED: return x[0] > y[0] ? log(0.75) : log(0.25);
      default: return 0.0;
    }
  }
};

}  // namespace mg
