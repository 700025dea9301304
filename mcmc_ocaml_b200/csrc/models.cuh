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
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

#include "../../include/mcmc_gpu.h"
#include "rng.cuh"

namespace mg {

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xFFF0000000000000ll); }
__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7FF8000000000000ll); }

// Dynamic plugins unroll their per-dimension loops only for small DMAX (state
// in registers); larger DMAX keep rolled loops over local-memory arrays.
#define MG_DYN_UNROLL(DMAX) ((DMAX) <= 8 ? (DMAX) : 1)

#ifndef MG_USER_EVAL
#define MG_USER_EVAL(kind, x, d, p, np) (mg::neg_inf())
#endif

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
template <class R>
__device__ __forceinline__ double draw_gaussian(R &r, double mu, double sigma) {
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
// stats.ml:89-91
template <class R>
__device__ __forceinline__ double draw_cauchy(R &r, double x0, double gamma) {
  const double p = r.uniform();
  return x0 + gamma * tan(3.14159265358979323846 * (p - 0.5));
}
// stats.ml:126-128
template <class R>
__device__ __forceinline__ double draw_uniform(R &r, double a, double b) {
  const double d = b - a;
  return a + d * r.uniform();
}
// mcmc.ml:187-196 (reflects at the bounds)
template <class R>
__device__ __forceinline__ double uniform_wrapping(R &r, double xmin, double xmax, double dx, double x) {
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

// Static plugins read their parameters from SHARED memory with volatile
// 16-byte loads.  On sm_100a a constant-bank operand is first loaded into a
// uniform register (LDCU); with ~85 parameters live in the step loop ptxas
// ran out of uniform registers and shuttled them through vector registers
// (ncu, profiles/r01_mh_ncu_summary.md: 75 R2UR + 61 IMAD.U32 + 48 LDCU of
// ~800 instructions per chain-step).  A broadcast LDS.128 per parameter pair
// replaces all of that; `volatile` keeps the loads inside the loop instead of
// being hoisted into 170 registers.
__device__ __forceinline__ double2 lds2(const double *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
// non-volatile form: the compiler may keep the pair in registers across steps (loop-invariant); used for the few
// parameters the register budget has room for (experiment switches MG_EXP_REG_PROP / MG_EXP_REG_MU)
__device__ __forceinline__ double2 lds2_hoistable(const double *p) {
  double2 v;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ double lds1(const double *p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

struct ZeroFn {
  struct Params {};
  static constexpr int kSmem = 0;
  template <int D>
  static __device__ __forceinline__ double eval(const Params &, const double *, const double (&)[D], int) { return 0.0; }
};

// MG_FN_GAUSS_CORR: logc - 1/2 |L (x - mu)|^2, L lower triangular.
// Parameter block (built on the host by GaussCorr<D>::pack): mu padded to an
// even count, then the rows of L each padded to an even count, then logc.
template <int D>
struct GaussCorr {
  static constexpr int kMu = (D + 1) & ~1;
  // rows have padded lengths 2,2,4,4,6,6,...: offset(i) = kMu + sum_{r<i} pad(r+1)
  static __host__ __device__ constexpr int pad_len(int i) { return (i + 2) & ~1; }
  static __host__ __device__ constexpr int off(int i) { int o = kMu; for (int r = 0; r < i; ++r) o += pad_len(r); return o; }
  static constexpr int kLogc = off(D);
  static constexpr int kSmem = kLogc + 2;
  struct Params { double s[kSmem]; };
  static void pack(const double *mu, const double *Lpacked, double logc, Params &p) {
    for (int k = 0; k < kSmem; ++k) p.s[k] = 0.0;
    for (int j = 0; j < D; ++j) p.s[j] = mu[j];
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) p.s[off(i) + j] = Lpacked[i * (i + 1) / 2 + j];
    p.s[kLogc] = logc;
  }
  template <int DD>
  static __device__ __forceinline__ double eval(const Params &, const double *s, const double (&x)[DD], int) {
    static_assert(DD == D, "GaussCorr: dimension mismatch");
#ifdef MG_EXP_NOLIKE   /* timing experiment only (tools/mh_ablation.sh): a log-density of D adds */
    { double q = lds1(s + kLogc); for (int j = 0; j < D; ++j) q = q - x[j]; return q; }
#endif
    double z[D];
#pragma unroll
    for (int j = 0; j < D; j += 2) {
#ifdef MG_EXP_REG_MU
      const double2 m = lds2_hoistable(s + j);
#else
      const double2 m = lds2(s + j);
#endif
      z[j] = x[j] - m.x;
      if (j + 1 < D) z[j + 1] = x[j + 1] - m.y;
    }
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double y = 0.0;
#pragma unroll
      for (int j = 0; j <= i; j += 2) {
        const double2 l = lds2(s + off(i) + j);
        y = (j == 0) ? l.x * z[0] : fma(l.x, z[j], y);
        if (j + 1 <= i) y = fma(l.y, z[j + 1], y);
      }
      q = fma(y, y, q);
    }
    return fma(-0.5, q, lds1(s + kLogc));
  }
};

// MG_PROP_BOX: x_i + random_between (-h_i) h_i  (bin/evidence_direct.ml:24-43)
template <int D>
struct BoxProp {
  // x_i + random_between (-h_i) h_i = x_i + (a + (b - a) u), evaluated with
  // m = 1 + u in [1, 2) as ONE fused multiply-add: x_i + fma(w_i, m, c_i),
  // (c_i, w_i) = (a - (b - a), b - a) pairs formed on the host (oracle: og_models.hpp).
  static constexpr int kSmem = 2 * D;
  struct Params { double s[kSmem]; };
  static constexpr bool kSymmetric = true;
  static constexpr bool kStaticDim = true;
  static constexpr int kDraws = D;  // uniforms consumed per proposal (fixed)
  static void pack(const double *h, Params &p) {
    for (int i = 0; i < D; ++i) { const double lo = -h[i], hi = h[i], w = hi - lo; p.s[2 * i] = lo - w; p.s[2 * i + 1] = w; }
  }
  template <int DD, class R>
  static __device__ __forceinline__ void propose(const Params &, const double *s, R &r, const double (&x)[DD],
                                                 double (&y)[DD], int) {
    static_assert(DD == D, "BoxProp: dimension mismatch");
#pragma unroll
    for (int i = 0; i < D; ++i) {
#ifdef MG_EXP_REG_PROP
      const double2 cw = lds2_hoistable(s + 2 * i);
#else
      const double2 cw = lds2(s + 2 * i);
#endif
      y[i] = x[i] + fma(cw.y, r.uniform12(), cw.x);
    }
  }
  template <int DD>
  static __device__ __forceinline__ double log_q(const Params &, const double *, const double (&)[DD],
                                                 const double (&)[DD], int) {
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
__device__ __forceinline__ double dyn_log_multi_gaussian(const double *mu, const double *sigma, const double *log_sigma,
                                                         const double (&x)[DMAX], int d) {
  // log sigma_i does not depend on x: it is evaluated once on the host (host_plugins.hpp appends it to the blob)
  double result = 0.0;  // stats.ml:103-108
#pragma unroll (DMAX <= 16 ? DMAX : 1)
  for (int i = 0; i < DMAX; ++i)
    if (i < d) result = result + log_gaussian_ls(__ldg(mu + i), __ldg(sigma + i), __ldg(log_sigma + i), x[i]);
  return result + 0.0;
}

struct DynFn {
  typedef DynFnParams Params;
  static constexpr int kSmem = 0;
  template <int DMAX>
  static __device__ __forceinline__ double raw(const Params &f, const double (&x)[DMAX], int d) {
    const double *p = f.p;
    switch (f.kind) {
      case MG_FN_ZERO: return 0.0;
      case MG_FN_CONST: return __ldg(p);
      case MG_FN_BOX_CLOSED: {  // bin/gaussian_cauchy_efficiency.ml:60-67
        // no short circuit: the 2 d bounds are independent loads and the comparisons combine bitwise
        // (same truth value; a short-circuit chain serialises the loads behind the predicates)
        int out = 0;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) { const double lo = __ldg(p + i), hi = __ldg(p + d + i); out |= (int)(x[i] < lo) | (int)(x[i] > hi); }
        const double inside = __ldg(p + 2 * d);
        return out ? neg_inf() : inside;
      }
      case MG_FN_BOX_OPEN: {  // test/nested_test.ml:24-28
        int in = 1;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) { const double lo = __ldg(p + i), hi = __ldg(p + d + i); in &= (int)(x[i] > lo) & (int)(x[i] < hi); }
        const double inside = __ldg(p + 2 * d);
        return in ? inside : neg_inf();
      }
      case MG_FN_GAUSS_DIAG: return dyn_log_multi_gaussian<DMAX>(p, p + d, p + f.np, x, d);
      case MG_FN_GAUSS_CORR: {
        const double *mu = p, *L = p + d;
        double z[DMAX];
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int j = 0; j < DMAX; ++j) z[j] = (j < d) ? x[j] - __ldg(mu + j) : 0.0;
        double q = 0.0;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i) {
          if (i < d) {
            double y = __ldg(L + i * (i + 1) / 2) * z[0];
#pragma unroll (DMAX <= 16 ? DMAX : 1)
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
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) { const double dx = x[i] - __ldg(p + i); s = s + dx * dx; }
        return log_gaussian_ls(__ldg(p + d), __ldg(p + d + 1), __ldg(p + f.np), sqrt(s));  // [np] = log sigma (host)
      }
      case MG_FN_GAUSS_MIX: {  // test/nested_test.ml:47-53
        const int K = (int)__ldg(p);
        double tot = 0.0;
        for (int k = 0; k < K; ++k) tot = tot + exp(dyn_log_multi_gaussian<DMAX>(p + 1 + k * d, p + 1 + K * d, p + f.np, x, d));
        return log(tot);
      }
    }
    // kinds >= MG_FN_USER: functions registered with mg_plugin_register_source; they
    // exist only in kernels compiled at run time (jit.cu), where MG_USER_EVAL is defined
    return MG_USER_EVAL(f.kind, x, d, p, f.np);
  }
  template <int DMAX>
  static __device__ __forceinline__ double eval(const Params &f, const double *, const double (&x)[DMAX], int d) {
    const double v = raw<DMAX>(f, x, d);
    return f.scale == 1.0 ? v : f.scale * v;
  }
};

struct DynPropParams {
  int32_t kind, dim;
  const double *p;  // device
  int64_t np;
};

// mcmc.ml:155-163 (the Mcmc copy of log_sum_logs: log (1 + exp), not log1p)
__device__ __forceinline__ double mcmc_log_sum_logs(double la, double lb) {
  if (la == neg_inf() && lb == neg_inf()) return neg_inf();
  if (la > lb) { const double lr = lb - la; return la + log(1.0 + exp(lr)); }
  const double lr = la - lb;
  return lb + log(1.0 + exp(lr));
}

struct DynProp {
  typedef DynPropParams Params;
  static constexpr bool kSymmetric = false;
  static constexpr bool kStaticDim = false;
  static constexpr int kDraws = -1;  // data dependent (rejection loops)
  static constexpr int kSmem = 0;
  // one basic (non-mixture) proposal kind with parameters at p
  template <int DMAX, class R>
  static __device__ __forceinline__ void propose_basic(int kind, const double *p, R &r, const double (&x)[DMAX],
                                                       double (&y)[DMAX], int d) {
    switch (kind) {
      case MG_PROP_BOX:
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) { const double a = -__ldg(p + i), b = __ldg(p + i), w = b - a; y[i] = x[i] + fma(w, r.uniform12(), a - w); }
        break;
      case MG_PROP_WRAP:
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) y[i] = uniform_wrapping(r, __ldg(p + i), __ldg(p + d + i), __ldg(p + 2 * d + i), x[i]);
        break;
      case MG_PROP_INDEP_GAUSS:
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) y[i] = draw_gaussian(r, __ldg(p + i), __ldg(p + d + i));
        break;
      case MG_PROP_LEFT_BIASED:  // test/mcmc_test.ml:66-70
        if (r.uniform() < 0.75) y[0] = x[0] - __ldg(p) * r.uniform();
        else y[0] = x[0] + __ldg(p) * r.uniform();
        break;
      case MG_PROP_ONE_SIDED:    // test/mcmc_test.ml:186-189
        y[0] = x[0] + __ldg(p) * (__ldg(p + 1) * r.uniform());
        break;
      case MG_PROP_DE: {         // differential_evolution_proposal, mcmc.ml:198-218
        const double mode_hop = __ldg(p);
        const uint64_t M = (uint64_t)__ldg(p + 1);
        const uint64_t i = r.below(M);                                   // :201-202 pick_samples: i, then j <> i
        uint64_t j;
        do { j = r.below(M); } while (j == i);
        double dd;
        if (mode_hop != 0.0 && r.uniform() < mode_hop) dd = 1.0;           // :209-210
        else dd = draw_gaussian(r, 0.0, 2.38 / sqrt(2.0 * (double)d));     // :212-213
        const double *xs = p + 2 + i * (uint64_t)d, *ys = p + 2 + j * (uint64_t)d;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int k = 0; k < DMAX; ++k)
          if (k < d) y[k] = x[k] + dd * (__ldg(ys + k) - __ldg(xs + k));   // :215-217
        break;
      }
    }
  }
  template <int DMAX>
  static __device__ __forceinline__ double log_q_basic(int kind, const double *p, const double (&x)[DMAX],
                                                       const double (&y)[DMAX], int d) {
    switch (kind) {
      case MG_PROP_INDEP_GAUSS: {  // test/mcmc_test.ml:123-126
        double s = 0.0;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < d) s = s + log_gaussian(__ldg(p + i), __ldg(p + d + i), y[i]);
        return s;
      }
      case MG_PROP_LEFT_BIASED: return x[0] > y[0] ? log(0.75) : log(0.25);
      case MG_PROP_ONE_SIDED: {    // test/mcmc_test.ml:190-199
        const double dd = __ldg(p) * (y[0] - x[0]);
        return (dd >= 0.0 && dd <= __ldg(p + 1)) ? 0.0 - log(__ldg(p + 1)) : neg_inf();
      }
      default: return 0.0;
    }
  }
  template <int DMAX, class R>
  static __device__ __forceinline__ void propose(const Params &f, const double *, R &r, const double (&x)[DMAX],
                                                 double (&y)[DMAX], int d) {
    const double *p = f.p;
    if (f.kind == MG_PROP_MIXTURE) {  // combine_jump_proposals, mcmc.ml:165-176
      const int K = (int)__ldg(p);
      double ptot = 0.0;
      { const double *q = p + 1; for (int c = 0; c < K; ++c) { ptot = ptot + __ldg(q); q += 3 + (int)__ldg(q + 2); } }
      double prob = r.uniform();
      const double *q = p + 1;
      for (int c = 0; c < K; ++c) {
        const double w = __ldg(q) / ptot;
        if (prob < w || c == K - 1) break;   // (the reference raises Failure if nothing is selected)
        prob = prob - w;
        q += 3 + (int)__ldg(q + 2);
      }
      propose_basic<DMAX, R>((int)__ldg(q + 1), q + 3, r, x, y, d);
    } else {
      propose_basic<DMAX, R>(f.kind, p, r, x, y, d);
    }
#pragma unroll (DMAX <= 16 ? DMAX : 1)
    for (int i = 0; i < DMAX; ++i)
      if (i >= d) y[i] = 0.0;
  }
  template <int DMAX>
  static __device__ __forceinline__ double log_q(const Params &f, const double *, const double (&x)[DMAX],
                                                 const double (&y)[DMAX], int d) {
    if (f.kind != MG_PROP_MIXTURE) return log_q_basic<DMAX>(f.kind, f.p, x, y, d);
    const double *p = f.p;            // mcmc.ml:177-184
    const int K = (int)__ldg(p);
    double ptot = 0.0;
    { const double *q = p + 1; for (int c = 0; c < K; ++c) { ptot = ptot + __ldg(q); q += 3 + (int)__ldg(q + 2); } }
    double log_jump = neg_inf();
    const double *q = p + 1;
    for (int c = 0; c < K; ++c) {
      const double log_local = log(__ldg(q) / ptot) + log_q_basic<DMAX>((int)__ldg(q + 1), q + 3, x, y, d);
      log_jump = mcmc_log_sum_logs(log_jump, log_local);
      q += 3 + (int)__ldg(q + 2);
    }
    return log_jump;
  }
};

}  // namespace mg
