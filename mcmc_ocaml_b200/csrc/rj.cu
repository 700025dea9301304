// rj.cu -- Mcmc.make_rjmcmc_sampler / rjmcmc_array (mcmc.ml:83-153): ensembles
// of independent two-model reversible-jump chains, one thread per chain.
//
// A chain's state is (model, x[Dmax], ll, lp) in registers.  Cross-model
// proposals are Interpolate_pdf draws.  The reference descends a kd-tree three
// times per cross-model step (draw into the target model, density of the drawn
// point, density of the current point for the reverse jump,
// test/mcmc_test.ml:175-178); here the drawn point's density comes with its
// draw and the current point's is kept until the point changes, so most
// cross-model steps descend once.  The descent keeps its box in shared memory
// (kdtree.cuh).  Traffic per step: 1 byte (model) or
// 8 (Dmax+2) bytes (full sample) per recorded chain-step plus ~16 B per tree
// level gathered through L2 -- latency-bound pointer chasing, not HBM-bound.
#include "common.cuh"
#include "host_plugins.hpp"
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

struct RjArgs {
  RjModelDev m[2];
  int64_t C, nbin, nskip, n;
  uint64_t chain_offset;
  CallKey key;
  int32_t Dm, DT;            // max model dim; scratch dim (max tree dim)
  uint8_t *out_model;        // [n][C] or null
  double *out_samples;       // [n][Dm+2][C] or null
  unsigned long long *counts;  // [5]: #A, #B, #accepted, #cross-model proposals, #cross-model accepted
  const double *start;         // [2][64]: the start points a0, b0
  int *fail;
};

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

template <int DMAX>
__global__ void __maxnreg__(MG_RJ_MAXNREG(DMAX))
rj_ensemble_kernel(const __grid_constant__ RjArgs a) {
  extern __shared__ double smem[];
  const KdScratch s = kd_scratch(smem, a.DT);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = c < a.C;
  unsigned na = 0, nb = 0, nacc = 0, ncross = 0, ncross_acc = 0;
  if (live) {
    const uint64_t g = a.chain_offset + (uint64_t)c;
    const int64_t C = a.C;
    const int F = a.Dm + 2;
    // rjmcmc_array mcmc.ml:121-128: fair coin for the initial model (F5a)
    Rng r0(a.key, P_RJ_INIT, g, 0);
    int model = (r0.uniform() < 0.5) ? 0 : 1;
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
      int pmodel;
      if (r.uniform() < cm.p) {               // :94,99 stay in the model
        pmodel = model;
        DynProp::propose<DMAX>(cm.prop, nullptr, r, x, y, cm.D);
      } else {                                // :97,102 jump into the other model
        pmodel = 1 - model;
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
      if (model == 0) ++na; else ++nb;
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
    na += __shfl_down_sync(0xffffffffu, na, off);
    nb += __shfl_down_sync(0xffffffffu, nb, off);
    nacc += __shfl_down_sync(0xffffffffu, nacc, off);
    ncross += __shfl_down_sync(0xffffffffu, ncross, off);
    ncross_acc += __shfl_down_sync(0xffffffffu, ncross_acc, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(a.counts + 0, (unsigned long long)na);
    atomicAdd(a.counts + 1, (unsigned long long)nb);
    atomicAdd(a.counts + 2, (unsigned long long)nacc);
    atomicAdd(a.counts + 3, (unsigned long long)ncross);
    atomicAdd(a.counts + 4, (unsigned long long)ncross_acc);
  }
}

}  // namespace mg

using namespace mg;

extern "C" int mg_rjmcmc_array(mg_ctx *ctx, const mg_rj_model *A, const mg_rj_model *B, const mg_rjmcmc_cfg *cfg,
                               const double *a0, const double *b0, uint8_t *out_model, double *out_samples,
                               int64_t out_counts[2]) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, A && B && cfg && a0 && b0 && out_counts, "rjmcmc_array: null argument");
  MG_REQUIRE(ctx, cfg->nchains >= 1 && cfg->nbin >= 0 && cfg->nskip >= 1 && cfg->n >= 0, "rjmcmc_array: bad nbin/nskip/n");
  // mcmc.ml:90 assert(pa +. pb -. 1.0 < sqrt epsilon_float)  -- one-sided, as coded
  if (!(A->p + B->p - 1.0 < sqrt(2.220446049250313e-16))) return set_err(ctx, MG_EFAIL, "Assert_failure mcmc.ml:90");
  const mg_rj_model *M[2] = {A, B};
  int rc;
  int DT = 0, Dm = 0;
  for (int k = 0; k < 2; ++k) {
    const int D = M[k]->like.dim;
    MG_REQUIRE(ctx, D >= 1 && D <= 64, "rjmcmc_array: dim must be in 1..64");
    if ((rc = validate_logfn(ctx, &M[k]->like, D, "log_likelihood"))) return rc;
    if ((rc = validate_logfn(ctx, &M[k]->prior, D, "log_prior"))) return rc;
    MG_REQUIRE(ctx, M[k]->like.kind < MG_FN_USER && M[k]->prior.kind < MG_FN_USER,
               "rjmcmc_array: run-time plugins are supported by mcmc_array and logfn_eval only");
    if ((rc = validate_proposal(ctx, &M[k]->prop, D))) return rc;
    if (M[k]->into.kind == MG_INTO_INTERP) {
      MG_REQUIRE(ctx, M[k]->into.tree != nullptr, "rjmcmc_array: interpolated jump without a tree");
      MG_REQUIRE(ctx, M[k]->into.tree->h.D == D, "rjmcmc_array: tree dimension != model dimension");
      DT = DT > D ? DT : D;
    } else if (M[k]->into.kind == MG_INTO_INDEP_GAUSS) {
      MG_REQUIRE(ctx, M[k]->into.params && M[k]->into.nparams == 2 * D, "rjmcmc_array: into-gaussian needs mu[D], sigma[D]");
    } else return set_err(ctx, MG_EINVAL, "rjmcmc_array: unknown into-proposal kind %d", M[k]->into.kind);
    Dm = Dm > D ? Dm : D;
  }
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int64_t C = cfg->nchains, n = cfg->n;
  const int F = Dm + 2;
  DevLogFn dl[2], dp[2]; DevProposal dj[2];
  DevBuf<double> dinto[2], d_samples, d_t;
  DevBuf<uint8_t> d_model;
  DevBuf<unsigned long long> d_counts;
  DevBuf<double> d_start;
  DevBuf<int> d_fail;
  RjArgs a{};
  std::vector<double> h_start(128, 0.0);
  for (int i = 0; i < A->like.dim; ++i) h_start[i] = a0[i];
  for (int i = 0; i < B->like.dim; ++i) h_start[64 + i] = b0[i];
  MG_CUDA(ctx, upload(d_start, h_start.data(), h_start.size(), s));
  MG_CUDA(ctx, d_counts.alloc(8, s));
  MG_CUDA(ctx, cudaMemsetAsync(d_counts.get(), 0, 8 * sizeof(unsigned long long), s));
  MG_CUDA(ctx, d_fail.alloc(1, s));
  MG_CUDA(ctx, cudaMemsetAsync(d_fail.get(), 0, sizeof(int), s));
  for (int k = 0; k < 2; ++k) {
    MG_CUDA(ctx, dl[k].upload_from(&M[k]->like, s));
    MG_CUDA(ctx, dp[k].upload_from(&M[k]->prior, s));
    MG_CUDA(ctx, dj[k].upload_from(&M[k]->prop, s));
    RjModelDev &m = a.m[k];
    m.like = dl[k].params; m.prior = dp[k].params; m.prop = dj[k].params;
    m.into_kind = M[k]->into.kind; m.nstop = M[k]->into.nstop; m.D = M[k]->like.dim;
    m.p = M[k]->p; m.log_p = log(M[k]->p);   // mcmc.ml:91, host libm like the reference
    m.into_p = nullptr;
    memset(&m.tree, 0, sizeof m.tree);
    if (m.into_kind == MG_INTO_INTERP) m.tree = M[k]->into.tree->view();
    else {
      MG_CUDA(ctx, upload(dinto[k], M[k]->into.params, (size_t)M[k]->into.nparams, s));
      m.into_p = dinto[k].get();
    }
  }
  if (out_model && n > 0) MG_CUDA(ctx, d_model.alloc((size_t)n * C, s));
  if (out_samples && n > 0) MG_CUDA(ctx, d_samples.alloc((size_t)n * F * C, s));
  a.C = C; a.nbin = cfg->nbin; a.nskip = cfg->nskip; a.n = n; a.chain_offset = cfg->chain_offset;
  a.key = next_key(ctx);
  a.Dm = Dm; a.DT = DT;
  a.out_model = d_model.get(); a.out_samples = d_samples.get();
  a.counts = d_counts.get(); a.start = d_start.get(); a.fail = d_fail.get();
  const int block = DT <= 32 ? 64 : 32;
  const size_t smem = kd_scratch_bytes(DT > 0 ? DT : 1, block);
  const unsigned grid = (unsigned)((C + block - 1) / block);
  time_begin(ctx);
#define MG_RJ_LAUNCH(DD)                                                                              \
  do {                                                                                                \
    if (smem > 48 * 1024)                                                                             \
      MG_CUDA(ctx, cudaFuncSetAttribute(rj_ensemble_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    rj_ensemble_kernel<DD><<<grid, block, smem, s>>>(a);                                              \
  } while (0)
  if (Dm <= 2) MG_RJ_LAUNCH(2);
  else if (Dm <= 4) MG_RJ_LAUNCH(4);
  else if (Dm <= 8) MG_RJ_LAUNCH(8);
  else if (Dm <= 16) MG_RJ_LAUNCH(16);
  else if (Dm <= 32) MG_RJ_LAUNCH(32);
  else MG_RJ_LAUNCH(64);
#undef MG_RJ_LAUNCH
  MG_CHECK_LAUNCH(ctx);
  time_end(ctx);
  if (out_model && n > 0) MG_CUDA(ctx, cudaMemcpyAsync(out_model, d_model.get(), (size_t)n * C, cudaMemcpyDeviceToHost, s));
  if (out_samples && n > 0)
    MG_CUDA(ctx, cudaMemcpyAsync(out_samples, d_samples.get(), sizeof(double) * (size_t)n * F * C, cudaMemcpyDeviceToHost, s));
  unsigned long long cnt[5];
  int h_fail = 0;
  MG_CUDA(ctx, cudaMemcpyAsync(cnt, d_counts.get(), sizeof cnt, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaMemcpyAsync(&h_fail, d_fail.get(), sizeof(int), cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  if (h_fail) return set_err(ctx, MG_EFAIL, "draw: empty tree");   // interpolate_pdf.ml:117,124
  out_counts[0] = (int64_t)cnt[0]; out_counts[1] = (int64_t)cnt[1];
  const int64_t steps = C * (cfg->nbin + (n > 0 ? (n - 1) * cfg->nskip : 0));
  ctx->naccept += (int64_t)cnt[2]; ctx->nreject += steps - (int64_t)cnt[2];
  ctx->rj_cross[0] = (int64_t)cnt[3]; ctx->rj_cross[1] = (int64_t)cnt[4];
  return MG_OK;
}

extern "C" int mg_rjmcmc_jump_counters(const mg_ctx *ctx, int64_t *proposed, int64_t *accepted) {
  if (!ctx) return MG_EINVAL;
  if (proposed) *proposed = ctx->rj_cross[0];
  if (accepted) *accepted = ctx->rj_cross[1];
  return MG_OK;
}
