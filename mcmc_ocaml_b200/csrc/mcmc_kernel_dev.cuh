// mcmc_kernel_dev.cuh -- device side of the fused Metropolis-Hastings ensemble
// kernel (no host includes: this header is also compiled at run time by NVRTC
// for user-registered plugins, see jit.cu).
#pragma once
#include "models.cuh"
#include "rng.cuh"

namespace mg {

#ifndef MG_MH_MAXNREG
#define MG_MH_MAXNREG(D) ((D) <= 10 ? 128 : 255)
#endif
#ifndef MG_MH_PIPELINE
#define MG_MH_PIPELINE 0
#endif
#ifndef MG_MH_BLOCK
#define MG_MH_BLOCK 32
#endif
constexpr int MH_BLOCK = MG_MH_BLOCK;  // 65,536 chains -> 2048 one-warp CTAs = 13.8 per SM on 148 SMs

template <class Like, class Prior, class Prop, int D>
struct MhArgs {
  typename Like::Params like;
  typename Prior::Params prior;
  typename Prop::Params prop;
  int32_t d, pad;
  int64_t C;
  uint64_t chain_offset;
  int64_t nbin, nskip, n;
  uint64_t t0;          // index of the first step of this launch (a run may be split into segments)
  int32_t record_first; // 1: slot 0 = the state after burn-in (mcmc.ml:66); 0: a continuation segment
  int32_t pad2;
  CallKey key;
  RoundKeys rk;     // key + r * W for the ten Philox rounds
  double *state;    // [D+2][C] in/out
  double *samples;  // [n][D+2][C] or null
  int32_t *accept;  // [C] accumulated, or null
  double *mom;      // [3][D+2][C] per-chain pivot, sum (v - pivot), sum (v - pivot)^2 of the recorded samples, or null
};

// Recording a sample: field i lives i * C doubles after field 0.  With the pitch of a field row in BYTES held as a
// 32-bit value, every field address is one IMAD.WIDE.U32 (pitch * i + base) instead of a 64-bit multiply and add
// per field (47 of the 59 instructions of a 12-field record were address arithmetic).  pitch32 == 0: the block is
// too wide for that (C * 8 * (D + 2) >= 2^32), 64-bit path.
template <int D>
__device__ __forceinline__ void store_sample(double *out, uint32_t pitch32, int64_t C, int dd, const double (&x)[D],
                                             double ll, double lp) {
#ifdef MG_EXP_NOSTORE  /* timing experiment only (tools/mh_ablation.sh): one field instead of D + 2 */
  __stcs(out, ll + lp + x[0]); return;
#endif
  if (pitch32) {
    char *b = reinterpret_cast<char *>(out);
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i < dd) __stcs(reinterpret_cast<double *>(b + (uint64_t)pitch32 * (uint32_t)i), x[i]);
    __stcs(reinterpret_cast<double *>(b + (uint64_t)pitch32 * (uint32_t)dd), ll);
    __stcs(reinterpret_cast<double *>(b + (uint64_t)pitch32 * (uint32_t)(dd + 1)), lp);
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i < dd) __stcs(out + (int64_t)i * C, x[i]);
    __stcs(out + (int64_t)dd * C, ll);
    __stcs(out + (int64_t)(dd + 1) * C, lp);
  }
}
__device__ __forceinline__ uint32_t sample_pitch32(int64_t C, int F) {
  return ((uint64_t)C * 8ull * (uint64_t)F < (1ull << 32)) ? (uint32_t)(C * 8) : 0u;
}

// mcmc.ml:37-56 make_mcmc_sampler: one step.  Returns 1 on acceptance.
template <class Like, class Prior, class Prop, int D, class RNG>
__device__ __forceinline__ int mh_step(const MhArgs<Like, Prior, Prop, D> &a, const double *sl, const double *sp,
                                       const double *sj, RNG &r, double (&x)[D], double &ll, double &lp) {
  const int dd = Prop::kStaticDim ? D : a.d;
  const double start_log_post = ll + lp;
  double y[D];
  Prop::template propose<D, RNG>(a.prop, sj, r, x, y, dd);
  const double proposed_like = Like::template eval<D>(a.like, sl, y, dd);
  const double proposed_prior = Prior::template eval<D>(a.prior, sp, y, dd);
  const double proposed_log_posterior = proposed_like + proposed_prior;
  double log_accept_prob = proposed_log_posterior - start_log_post;
  if (!Prop::kSymmetric) {
    const double log_forward_jump = Prop::template log_q<D>(a.prop, sj, x, y, dd);
    const double log_backward_jump = Prop::template log_q<D>(a.prop, sj, y, x, dd);
    log_accept_prob = log_accept_prob + log_backward_jump - log_forward_jump;
  }
  // log (Random.float 1.0) < log_accept_prob, strict (mcmc.ml:47).  NaN rejects.
  const bool acc = log_u_less_than(r.uniform(), log_accept_prob);
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = acc ? y[i] : x[i];
  ll = acc ? proposed_like : ll;
  lp = acc ? proposed_prior : lp;
  return acc ? 1 : 0;
}

// kStaticDim: the run-time dimension equals D (static plugins), so every
// `i < d` guard folds away.  Register budget: 65,536 chains need 13.8 warps
// per SM to be resident at once; the register file is split per scheduler (16K each), so <= 128 registers keep 4 warps on each
// (the balanced sampler holds exactly 3 per scheduler and takes 160, mcmc_balanced.cuh).
template <class Like, class Prior, class Prop, int D>
__device__ __forceinline__ void mh_ensemble_body(const MhArgs<Like, Prior, Prop, D> &a) {
  int64_t c = (int64_t)blockIdx.x * MH_BLOCK + threadIdx.x;
  const bool live = c < a.C;
  if (!live) c = a.C - 1;  // keep the whole CTA alive for the staging barrier; results discarded
  const uint64_t g = a.chain_offset + (uint64_t)c;
  const int64_t C = a.C;
  const int dd = Prop::kStaticDim ? D : a.d;
  const int F = dd + 2;
  // stage the static plugins' parameter blocks in shared memory (see models.cuh)
  __shared__ __align__(16) double smem_params[Like::kSmem + Prior::kSmem + Prop::kSmem + 2];
  double *sl = smem_params, *sp = sl + Like::kSmem, *sj = sp + Prior::kSmem;
  if (Like::kSmem + Prior::kSmem + Prop::kSmem > 0) {
    const double *gl = reinterpret_cast<const double *>(&a.like);
    const double *gp = reinterpret_cast<const double *>(&a.prior);
    const double *gj = reinterpret_cast<const double *>(&a.prop);
    for (int k = threadIdx.x; k < Like::kSmem; k += MH_BLOCK) sl[k] = gl[k];
    for (int k = threadIdx.x; k < Prior::kSmem; k += MH_BLOCK) sp[k] = gp[k];
    for (int k = threadIdx.x; k < Prop::kSmem; k += MH_BLOCK) sj[k] = gj[k];
    __syncthreads();
  }
  double x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = (i < dd) ? a.state[(int64_t)i * C + c] : 0.0;
  // mcmc.ml:59-61: the start point is evaluated, not trusted
  double ll = Like::template eval<D>(a.like, sl, x, dd);
  double lp = Prior::template eval<D>(a.prior, sp, x, dd);
  int nacc = 0;
  uint64_t t = a.t0;
  // Fixed-draw proposals: the uniforms of step t+1 are generated while step t
  // computes (software pipelining across the loop edge, which the compiler
  // cannot do by itself).  Data-dependent proposals draw on demand.
  constexpr bool kPipe = (Prop::kDraws >= 0) && MG_MH_PIPELINE;
  constexpr int kNU = kPipe ? Prop::kDraws + 1 : 1;
  RngBuf<kNU> cur;
  if (kPipe) cur.fill(a.key, P_MH, g, a.t0);
  auto step = [&]() -> int {
    int r_acc;
    if constexpr (kPipe) {
      RngBuf<kNU> nxt;
      nxt.fill(a.key, P_MH, g, t + 1);
      r_acc = mh_step<Like, Prior, Prop, D>(a, sl, sp, sj, cur, x, ll, lp);
      cur = nxt;
    } else {
      Rng r(a.key, P_MH, g, t, &a.rk);
      r_acc = mh_step<Like, Prior, Prop, D>(a, sl, sp, sj, r, x, ll, lp);
    }
    ++t;
    return r_acc;
  };
  for (int64_t i = 0; i < a.nbin; ++i) nacc += step();  // :63-65
  double *out = (a.samples && live) ? a.samples + c : nullptr;
  const int64_t sample_stride = (int64_t)F * C;
  const uint32_t pitch32 = sample_pitch32(C, F);
  auto record = [&]() {
    if (out) {
      store_sample<D>(out, pitch32, C, dd, x, ll, lp);
      out += sample_stride;
    }
  };
  if (a.n > 0 && a.record_first) record();  // :66 slot 0 = state after burn-in
  for (int64_t s = 1; s < a.n; ++s) {  // :67-71
    for (int64_t k = 0; k < a.nskip; ++k) nacc += step();
    record();
  }
  if (!live) return;
#pragma unroll
  for (int i = 0; i < D; ++i)
    if (i < dd) a.state[(int64_t)i * C + c] = x[i];
  a.state[(int64_t)dd * C + c] = ll;
  a.state[(int64_t)(dd + 1) * C + c] = lp;
  if (a.accept) a.accept[c] += nacc;
}


template <class Like, class Prior, class Prop, int D>
__global__ void __maxnreg__(MG_MH_MAXNREG(D))
mh_ensemble_kernel(const __grid_constant__ MhArgs<Like, Prior, Prop, D> a) {
  mh_ensemble_body<Like, Prior, Prop, D>(a);
}

}  // namespace mg
