// rj_kernel_dev.cuh -- device side of the reversible-jump ensemble kernel (Mcmc.make_rjmcmc_sampler / rjmcmc_array,
// mcmc.ml:83-153).  No host includes: this header is also compiled at run time by NVRTC when a model uses a
// user-registered log-density (jit.cu).
#pragma once
#include "kdtree.cuh"
#include "models.cuh"

namespace mg {

struct RjModelDev {
  DynFnParams like, prior;
  DynPropParams prop;
  KdView tree;
  const double *into_p;  // MG_INTO_INDEP_GAUSS: mu[D], sigma[D]
  double p, log_p;
  int32_t into_kind, nstop, D, pad;
};

// NM = capacity of the model table.  NM = 2 is the reference's sum type (A | B, mcmc.ml:83-87) and the kernel of
// mg_rjmcmc_array; NM = MG_RJ_MAX_MODELS serves mg_rjmcmc_array_k, the k-model extension, where
// K <= NM models are live.
template <int NM>
struct RjArgsT {
  RjModelDev m[NM];
  int64_t C, nbin, nskip, n;
  uint64_t chain_offset;
  CallKey key;
  int32_t Dm, DT;            // max model dim; scratch dim (max tree dim)
  int32_t K, pad;            // live models (2 when NM == 2)
  uint8_t *out_model;        // [n][C] or null
  double *out_samples;       // [n][Dm+2][C] or null
  unsigned long long *counts;  // [NM + 3]: samples per model, then #accepted, #cross-model proposals, #cross-model accepted
  const double *start;         // [NM][64]: the start points (a0, b0, ...)
  int *fail;
};
using RjArgs = RjArgsT<2>;

// *lq / *lq_known: Interp.draw leaves the cell's box in the scratch.  When the drawn point lies strictly inside
// it, Interp.jump_prob of that point descends to the same cell (at every ancestor the point is inside the child on
// the path and, where the path goes right, strictly beyond the split), so its value -- count / (volume * N), the
// same expression as kd_jump_prob -- is taken here and the second descent is skipped.  A point on the box
// boundary (u = 0, or a degenerate cell) is left to the full descent.
template <int DMAX>
__device__ __forceinline__ bool rj_draw_into(const RjModelDev &m, const KdScratch &s, Rng &r, double (&y)[DMAX],
                                             double *lq, bool *lq_known) {
  *lq_known = false;
  if (m.into_kind == MG_INTO_INTERP) {
    int32_t node;
    if (!kd_draw(m.tree, s, m.nstop, r, &node)) return false;
    bool inside = true;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
    for (int i = 0; i < DMAX; ++i) {
      y[i] = (i < m.D) ? s.Q(i) : 0.0;
      if (i < m.D) inside = inside && (s.LO(i) < y[i]) && (y[i] < s.HI(i));
    }
    if (inside) {
      const double nobjs = (double)__ldg(m.tree.count + node);
      const double v = kd_cell_volume(s, m.tree.D);
      *lq = log(nobjs / (v * (double)m.tree.N));
      *lq_known = true;
    }
  } else {
#pragma unroll (DMAX <= 16 ? DMAX : 1)
    for (int i = 0; i < DMAX; ++i)
      y[i] = (i < m.D) ? draw_gaussian(r, __ldg(m.into_p + i), __ldg(m.into_p + m.D + i)) : 0.0;
  }
  return true;
}

// ljpintoa / ljpintob: log density of proposing `to` into model m
template <int DMAX>
__device__ __forceinline__ double rj_log_into(const RjModelDev &m, const KdScratch &s, const double (&to)[DMAX]) {
  if (m.into_kind == MG_INTO_INTERP) {
#pragma unroll (DMAX <= 16 ? DMAX : 1)
    for (int i = 0; i < DMAX; ++i)
      if (i < m.D) s.Q(i) = to[i];
    return log(kd_jump_prob(m.tree, s, m.nstop, nullptr));  // test/mcmc_test.ml:177-178
  }
  double acc = 0.0;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
  for (int i = 0; i < DMAX; ++i)
    if (i < m.D) acc = acc + log_gaussian(__ldg(m.into_p + i), __ldg(m.into_p + m.D + i), to[i]);
  return acc;
}

// Register cap: the step is bound by the latency of dependent tree-node loads, so resident warps count for more than
// a few spilled values (tools/rj_regs_sweep.sh, config 5 (2,4)-D: 143 regs 1.35e9, 120: 1.69e9, 92: 1.95e9, 80: 2.07e9,
// 64: 2.21e9 chain-steps/s).
#ifndef MG_RJ_MAXNREG
#define MG_RJ_MAXNREG(DMAX) ((DMAX) <= 8 ? 64 : ((DMAX) <= 16 ? 96 : 255))
#endif

// Model choice of a step.  Two models: mcmc.ml:92-102 -- stay with probability p of the current model, else the
// other one.  K models (extension): the same uniform walks the model priors cyclically from the current model
// (current, current + 1, ..., mod K) and takes the first whose cumulative probability exceeds it, the last one taking
// the remainder; with K = 2 that is the reference's rule, draw for draw.  The jump densities keep the reference's
// form: log p_target + log q_into_target (mcmc.ml:103-112).
template <int NM>
__device__ __forceinline__ int rj_pick_model(const RjArgsT<NM> &a, int model, double u) {
  if (NM == 2) return (u < a.m[model].p) ? model : 1 - model;
  double cum = 0.0;
  int j = model;
  for (int k = 0; k < a.K - 1; ++k) {
    cum = cum + a.m[j].p;
    if (u < cum) return j;
    j = (j + 1 == a.K) ? 0 : j + 1;
  }
  return j;
}

template <int DMAX, int NM = 2>
__device__ __forceinline__ void rj_ensemble_body(const RjArgsT<NM> &a) {
  extern __shared__ double smem[];
  const KdScratch s = kd_scratch(smem, a.DT);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = c < a.C;
  unsigned nmod[NM];
#pragma unroll
  for (int k = 0; k < NM; ++k) nmod[k] = 0;
  unsigned nacc = 0, ncross = 0, ncross_acc = 0;
  if (live) {
    const uint64_t g = a.chain_offset + (uint64_t)c;
    const int64_t C = a.C;
    const int F = a.Dm + 2;
    // rjmcmc_array mcmc.ml:121-128: fair coin for the initial model (F5a); K models: uniform over the K
    Rng r0(a.key, P_RJ_INIT, g, 0);
    int model;
    if (NM == 2) model = (r0.uniform() < 0.5) ? 0 : 1;
    else { model = (int)(r0.uniform() * (double)a.K); if (model >= a.K) model = a.K - 1; }
    double x[DMAX], y[DMAX];
#pragma unroll (DMAX <= 16 ? DMAX : 1)
    for (int i = 0; i < DMAX; ++i) x[i] = 0.0;
    const double *start = a.start + (size_t)model * 64;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
    for (int i = 0; i < DMAX; ++i)
      if (i < a.m[model].D) x[i] = start[i];
    double ll = DynFn::eval<DMAX>(a.m[model].like, nullptr, x, a.m[model].D);
    double lp = DynFn::eval<DMAX>(a.m[model].prior, nullptr, x, a.m[model].D) + a.m[model].log_p;  // :128
    uint64_t t = 0;
    bool bad = false;
    double lq_x = 0.0;          // log of the jump-in probability of the current point in its own model
    bool lq_x_valid = false;
    auto step = [&]() {
      Rng r(a.key, P_RJ, g, t);
      ++t;
      double fwd_lq = 0.0;
      bool fwd_known = false;
      const double start_log_post = ll + lp;
      const RjModelDev &cm = a.m[model];
      const int pmodel = rj_pick_model<NM>(a, model, r.uniform());
      if (pmodel == model) {                  // :94,99 stay in the model
        DynProp::propose<DMAX>(cm.prop, nullptr, r, x, y, cm.D);
      } else {                                // :97,102 jump into the other model
        if (!rj_draw_into<DMAX>(a.m[pmodel], s, r, y, &fwd_lq, &fwd_known)) { bad = true; return; }
      }
      const RjModelDev &pm = a.m[pmodel];
      const double proposed_like = DynFn::eval<DMAX>(pm.like, nullptr, y, pm.D);                 // :113-115
      const double proposed_prior = pm.log_p + DynFn::eval<DMAX>(pm.prior, nullptr, y, pm.D);    // :116-118
      const double proposed_log_posterior = proposed_like + proposed_prior;
      double log_forward_jump, log_backward_jump;                                               // :103-112
      if (pmodel == model) {
        log_forward_jump = pm.log_p + DynProp::log_q<DMAX>(pm.prop, nullptr, x, y, pm.D);
        log_backward_jump = cm.log_p + DynProp::log_q<DMAX>(cm.prop, nullptr, y, x, cm.D);
      } else {
        if (!fwd_known) fwd_lq = rj_log_into<DMAX>(pm, s, y);
        // log (jump into the current model at x) is a pure function of x: kept until x changes
        if (!lq_x_valid) { lq_x = rj_log_into<DMAX>(cm, s, x); lq_x_valid = true; }
        log_forward_jump = pm.log_p + fwd_lq;
        log_backward_jump = cm.log_p + lq_x;
      }
      const double log_accept_prob =
          proposed_log_posterior - start_log_post + log_backward_jump - log_forward_jump;
      if (pmodel != model) ++ncross;
      if (log_u_less_than(r.uniform(), log_accept_prob)) {
        if (pmodel != model) { lq_x = fwd_lq; lq_x_valid = true; ++ncross_acc; }   // the new point's own jump-in probability
        else lq_x_valid = false;
        model = pmodel;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i) x[i] = y[i];
        ll = proposed_like; lp = proposed_prior; ++nacc;
      }
    };
    auto record = [&](int64_t smp) {
#pragma unroll
      for (int k = 0; k < NM; ++k) nmod[k] += (model == k) ? 1u : 0u;
      if (a.out_model) a.out_model[smp * C + c] = (uint8_t)model;
      if (a.out_samples) {
        double *o = a.out_samples + smp * (int64_t)F * C + c;
#pragma unroll (DMAX <= 16 ? DMAX : 1)
        for (int i = 0; i < DMAX; ++i)
          if (i < a.Dm) __stcs(o + (int64_t)i * C, x[i]);
        __stcs(o + (int64_t)a.Dm * C, ll);
        __stcs(o + (int64_t)(a.Dm + 1) * C, lp);
      }
    };
    for (int64_t i = 0; i < a.nbin && !bad; ++i) step();   // :129-131
    if (a.n > 0) record(0);
    for (int64_t smp = 1; smp < a.n && !bad; ++smp) {      // :133-138
      for (int64_t k = 0; k < a.nskip && !bad; ++k) step();
      record(smp);
    }
    if (bad) *a.fail = 1;
  }
  // rjmcmc_model_counts (mcmc.ml:141-149): warp-reduce, one atomic per warp
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < NM; ++k) nmod[k] += __shfl_down_sync(0xffffffffu, nmod[k], off);
    nacc += __shfl_down_sync(0xffffffffu, nacc, off);
    ncross += __shfl_down_sync(0xffffffffu, ncross, off);
    ncross_acc += __shfl_down_sync(0xffffffffu, ncross_acc, off);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < NM; ++k) atomicAdd(a.counts + k, (unsigned long long)nmod[k]);
    atomicAdd(a.counts + NM + 0, (unsigned long long)nacc);
    atomicAdd(a.counts + NM + 1, (unsigned long long)ncross);
    atomicAdd(a.counts + NM + 2, (unsigned long long)ncross_acc);
  }
}

}  // namespace mg
