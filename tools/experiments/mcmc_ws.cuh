// RECORDED NEGATIVE EXPERIMENT (round 1), kept out of the product build: a warp-specialised variant of the MH sampler
// (producer warps generate the Philox uniforms into shared memory, consumer warps run the float64 step).  Slower than
// the single-role kernel on B200 (profiles/r01_mh_ncu_summary.md); not compiled into libmcmcgpu.so.
// mcmc_ws.cuh -- warp-specialised form of the Metropolis-Hastings ensemble
// kernel for the static plugins (GAUSS_CORR likelihood, flat prior, box
// proposal).  Same arithmetic, same Philox stream and same results as
// mh_ensemble_kernel (mcmc_kernel.cuh); only the mapping to warps changes.
//
// Why.  ncu on the single-role kernel (profiles/r01_mh_ncu_summary.md): 4 warps
// per scheduler, all walking the same phases in step (Philox: ALU + IMAD
// pipes; likelihood: FP64 pipe), 63 % issue utilisation, every pipe below
// 40 %.  Chains are the only parallelism (65,536 of them = 3.5 warps per
// scheduler), so the work of ONE chain is split over TWO warps that run
// different pipes at the same time:
//   producer warp (warpgroup 0): Philox4x32-10, uniform -> proposal offset
//       a_i + w_i * u_i for every coordinate, log u for the accept test;
//       integer pipes + a little FP64; needs ~40 registers;
//   consumer warp (warpgroup 1): y = x + offset, triangular whitening
//       product, accept / reject, streaming stores; FP64 pipe; ~88 registers.
// They are coupled through a 2-stage ring in shared memory guarded by
// mbarriers (full / empty per stage, 2 steps per stage).  A CTA is 4
// producer + 4 consumer warps = 128 chains; `setmaxnreg` moves registers from
// the producer warpgroup to the consumer warpgroup so that 4 CTAs (8 warps per
// scheduler instead of 4) stay resident per SM.
#pragma once
#include "mcmc_kernel.cuh"

namespace mg {

constexpr int WS_PAIRS = 4;            // producer/consumer warp pairs per CTA
constexpr int WS_BLOCK = WS_PAIRS * 64;
constexpr int WS_SPS = 2;              // steps per stage
constexpr int WS_NSTAGE = 2;
constexpr int WS_SLOTS = 6;            // double2 slots per step: 10 offsets + log u + pad

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  unsigned spins = 0;
  while (!done) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 28)) __trap();  // a protocol error must not hang the GPU
  }
}
__device__ __forceinline__ void sts2(double *p, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(a), "d"(b) : "memory");
}

template <int D>
struct WsSmem {
  static constexpr int kParams = GaussCorr<D>::kSmem + BoxProp<D>::kSmem + 2;
  static constexpr int kRingDoubles = WS_PAIRS * WS_NSTAGE * WS_SPS * WS_SLOTS * 32 * 2;
  static constexpr size_t bytes = sizeof(double) * (kParams + kRingDoubles) + sizeof(uint64_t) * WS_PAIRS * WS_NSTAGE * 2;
};

template <int D>
__global__ void __launch_bounds__(WS_BLOCK, 4)
mh_ws_kernel(const __grid_constant__ MhArgs<GaussCorr<D>, ZeroFn, BoxProp<D>, D> a) {
  static_assert(D <= 10, "ring slots are sized for D <= 10");
  extern __shared__ __align__(16) double smem[];
  double *sl = smem;                              // GaussCorr params
  double *sj = sl + GaussCorr<D>::kSmem;          // BoxProp params
  double *ring = smem + WsSmem<D>::kParams;
  uint64_t *bars = reinterpret_cast<uint64_t *>(ring + WsSmem<D>::kRingDoubles);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp & (WS_PAIRS - 1);
  const bool producer = warp < WS_PAIRS;          // warpgroup 0
  uint64_t *full = bars + pair * WS_NSTAGE * 2, *empty = full + WS_NSTAGE;
  {
    const double *gl = reinterpret_cast<const double *>(&a.like);
    const double *gj = reinterpret_cast<const double *>(&a.prop);
    for (int k = threadIdx.x; k < GaussCorr<D>::kSmem; k += WS_BLOCK) sl[k] = gl[k];
    for (int k = threadIdx.x; k < BoxProp<D>::kSmem; k += WS_BLOCK) sj[k] = gj[k];
    if (threadIdx.x < WS_PAIRS * WS_NSTAGE * 2) mbar_init(bars + threadIdx.x, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }
  int64_t c = (int64_t)blockIdx.x * (WS_PAIRS * 32) + pair * 32 + lane;
  const bool live = c < a.C;
  if (!live) c = a.C - 1;
  const uint64_t g = a.chain_offset + (uint64_t)c;
  const int64_t C = a.C;
  const int64_t total = a.nbin + (a.n > 0 ? (a.n - 1) * a.nskip : 0);
  double2 *myring = reinterpret_cast<double2 *>(ring) + (size_t)pair * WS_NSTAGE * WS_SPS * WS_SLOTS * 32 + lane;

  if (producer) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    int64_t t = 0;                                 // steps done in this launch; the RNG step index is t0 + t
    for (int64_t k = 0; t < total; ++k) {
      const int s = (int)(k % WS_NSTAGE);
      mbar_wait(empty + s, (unsigned)(((k / WS_NSTAGE) & 1) ^ 1));
      double2 *dst = myring + (size_t)s * WS_SPS * WS_SLOTS * 32;
#pragma unroll
      for (int q = 0; q < WS_SPS; ++q) {
        if (t < total) {
          Rng r(a.key, P_MH, g, a.t0 + (uint64_t)t, &a.rk);
          double2 *slot = dst + q * WS_SLOTS * 32;
#pragma unroll
          for (int p = 0; p < 5; ++p) {       // BoxProp offset fma(w_i, 1 + u_i, c_i) (bin/evidence_direct.ml:24-25)
            double o0 = 0.0, o1 = 0.0;
            if (2 * p < D) { const double2 cw = lds2(sj + 4 * p); o0 = fma(cw.y, r.uniform12(), cw.x); }
            if (2 * p + 1 < D) { const double2 cw = lds2(sj + 4 * p + 2); o1 = fma(cw.y, r.uniform12(), cw.x); }
            sts2(reinterpret_cast<double *>(slot + p * 32), o0, o1);
          }
          const double logu = log(r.uniform());   // log (Random.float 1.0), mcmc.ml:47
          sts2(reinterpret_cast<double *>(slot + 5 * 32), logu, 0.0);
          ++t;
        }
      }
      mbar_arrive(full + s);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    double x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = a.state[(int64_t)i * C + c];
    ZeroFn::Params zp;
    double ll = GaussCorr<D>::template eval<D>(a.like, sl, x, D);   // mcmc.ml:59-61
    double lp = ZeroFn::eval<D>(zp, nullptr, x, D);
    int nacc = 0;
    int64_t t = 0, k = 0;
    int q = WS_SPS;                                // position inside the current stage
    int s = 0;
    const double2 *src = myring;
    auto step = [&]() {
      if (q == WS_SPS) {                           // next stage
        s = (int)(k % WS_NSTAGE);
        mbar_wait(full + s, (unsigned)((k / WS_NSTAGE) & 1));
        src = myring + (size_t)s * WS_SPS * WS_SLOTS * 32;
        q = 0;
      }
      double off[12];
#pragma unroll
      for (int p = 0; p < WS_SLOTS; ++p) {
        const double2 v = lds2(reinterpret_cast<const double *>(src + (q * WS_SLOTS + p) * 32));
        off[2 * p] = v.x; off[2 * p + 1] = v.y;
      }
      ++q; ++t;
      if (q == WS_SPS || t == total) { mbar_arrive(empty + s); ++k; q = WS_SPS; }
      // make_mcmc_sampler (mcmc.ml:37-56)
      const double start_log_post = ll + lp;
      double y[D];
#pragma unroll
      for (int i = 0; i < D; ++i) y[i] = x[i] + off[i];
      const double proposed_like = GaussCorr<D>::template eval<D>(a.like, sl, y, D);
      const double proposed_prior = ZeroFn::eval<D>(zp, nullptr, y, D);
      const double proposed_log_posterior = proposed_like + proposed_prior;
      const double log_accept_prob = proposed_log_posterior - start_log_post;
      const bool acc = off[10] < log_accept_prob;
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = acc ? y[i] : x[i];
      ll = acc ? proposed_like : ll;
      lp = acc ? proposed_prior : lp;
      nacc += acc ? 1 : 0;
    };
    for (int64_t i = 0; i < a.nbin; ++i) step();   // :63-65
    double *out = (a.samples && live) ? a.samples + c : nullptr;
    const int64_t sample_stride = (int64_t)(D + 2) * C;
    auto record = [&]() {
      if (out) {
#pragma unroll
        for (int i = 0; i < D; ++i) __stcs(out + (int64_t)i * C, x[i]);
        __stcs(out + (int64_t)D * C, ll);
        __stcs(out + (int64_t)(D + 1) * C, lp);
        out += sample_stride;
      }
    };
    if (a.n > 0 && a.record_first) record();       // :66
    for (int64_t smp = 1; smp < a.n; ++smp) {      // :67-71
      for (int64_t kk = 0; kk < a.nskip; ++kk) step();
      record();
    }
    if (live) {
#pragma unroll
      for (int i = 0; i < D; ++i) a.state[(int64_t)i * C + c] = x[i];
      a.state[(int64_t)D * C + c] = ll;
      a.state[(int64_t)(D + 1) * C + c] = lp;
      if (a.accept) a.accept[c] += nacc;
    }
  }
}

template <int D>
static int launch_mh_ws(mg_ctx *ctx, const MhArgs<GaussCorr<D>, ZeroFn, BoxProp<D>, D> &a) {
  const int64_t grid = (a.C + WS_PAIRS * 32 - 1) / (WS_PAIRS * 32);
  const size_t smem = WsSmem<D>::bytes;
  MG_CUDA(ctx, cudaFuncSetAttribute(mh_ws_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  time_begin(ctx);
  mh_ws_kernel<D><<<(unsigned)grid, WS_BLOCK, smem, ctx->stream>>>(a);
  MG_CHECK_LAUNCH(ctx);
  time_end(ctx);
  return MG_OK;
}

}  // namespace mg
