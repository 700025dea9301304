// oracle.cpp -- TEST INFRASTRUCTURE.  CPU restatement of the sampling-and-
// evidence path of farr/mcmc-ocaml (mcmc.ml, kd_tree.ml, interpolate_pdf.ml,
// evidence.ml, nested.ml, stats.ml).  It is the parity checker for the CUDA
// path and the "port" CPU baseline of bench.py; nothing in the product imports,
// links or calls it (tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs only).
//
// Parity status: the reference cannot be compiled here (no OCaml toolchain,
// SURVEY.md F1) and ships almost no golden vectors (F8).  This restatement is
// pinned against every exact expectation the reference's tests hold
// (tests/test_oracle_*.py: stats_test.ml goldens, log_lognormal, the
// analytic harmonic-mean target, kd-tree invariants and depth, ratio 4.0,
// priors 0.1/0.9, nested 1 and 4, ...).  Autocorrelation: parity unpinned (F6).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -mfma).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "og_models.hpp"
#include "og_rng.hpp"
#include "og_tree.hpp"

using namespace og;

static thread_local std::string g_err;
static int fail(int code, const std::string &m) { g_err = m; return code; }

// Diagnostic for the parity tests: the smallest |log u - log_accept_prob| over the accept tests (mcmc.ml:47)
// since the last reset.  A GPU chain may legitimately leave the oracle's only at a step whose margin is a
// near-tie (the two libms differ in the last ulp of a transcendental); the tests check exactly that.
static thread_local double tl_margin = HUGE_VAL;
static inline void note_margin(double log_u, double log_accept_prob) {
  const double m = std::fabs(log_u - log_accept_prob);
  if (m < tl_margin) tl_margin = m;   // NaN (inf - inf) never compares smaller: not a tie
}

// ===========================================================================
// mcmc.ml
// ===========================================================================

// mcmc.ml:37-56 make_mcmc_sampler: one Metropolis-Hastings step on a flat
// float-array state.  `x` (D doubles), `ll`, `lp` are updated in place on
// acceptance; on rejection the sample is returned unchanged (:55).
template <class LL, class LP, class JP, class LJP>
static inline bool mh_step(Rng &r, int D, double *x, double &ll, double &lp, double *prop,
                           LL &&log_likelihood, LP &&log_prior, JP &&jump_proposal,
                           LJP &&log_jump_prob) {
  double start_log_post = ll + lp;
  jump_proposal(r, x, prop);
  double proposed_like = log_likelihood(prop), proposed_prior = log_prior(prop);
  double proposed_log_posterior = proposed_like + proposed_prior;
  double log_forward_jump = log_jump_prob(x, prop), log_backward_jump = log_jump_prob(prop, x);
  double log_accept_prob =
      proposed_log_posterior - start_log_post + log_backward_jump - log_forward_jump;
  const double log_u = std::log(r.uniform());
  note_margin(log_u, log_accept_prob);
  if (log_u < log_accept_prob) {  // strict <, :47
    for (int i = 0; i < D; ++i) x[i] = prop[i];
    ll = proposed_like; lp = proposed_prior;
    return true;
  }
  return false;
}

// mcmc.ml:58-72 mcmc_array for one chain with global id g.  Samples are
// written step-major: out[(s*(D+2) + f)*C + c].
static void mcmc_chain(const CallKey &ck, const LogFn &like, const LogFn &prior,
                       const Proposal &prop, const mg_mcmc_cfg &cfg, const double *start,
                       int64_t c, double *out, int64_t *acc, int64_t *rej, double *margin = nullptr) {
  const int D = cfg.dim; const int64_t C = cfg.nchains; const int F = D + 2;
  tl_margin = HUGE_VAL;
  const uint64_t g = cfg.chain_offset + (uint64_t)c;
  std::vector<double> x(start, start + D), p(D);
  double ll = like(x.data()), lp = prior(x.data());  // :59-61
  int64_t na = 0, nr = 0, t = 0;
  auto L = [&](const double *y) { return like(y); };
  auto P = [&](const double *y) { return prior(y); };
  auto J = [&](Rng &r, const double *a, double *b) { prop.propose(r, a, b); };
  auto Q = [&](const double *a, const double *b) { return prop.log_q(a, b); };
  auto record = [&](int64_t s) {
    if (margin) { margin[s * C + c] = tl_margin; tl_margin = HUGE_VAL; }   // min over the steps leading to slot s
    if (!out) return;
    for (int f = 0; f < D; ++f) out[(s * F + f) * C + c] = x[f];
    out[(s * F + D) * C + c] = ll; out[(s * F + D + 1) * C + c] = lp;
  };
  for (int64_t i = 0; i < cfg.nbin; ++i, ++t) {  // :63-65
    Rng r(ck, P_MH, g, (uint64_t)t);
    (mh_step(r, D, x.data(), ll, lp, p.data(), L, P, J, Q) ? na : nr)++;
  }
  if (cfg.n > 0) record(0);  // :66
  for (int64_t i = 1; i <= (cfg.n - 1) * cfg.nskip; ++i, ++t) {  // :67-71
    Rng r(ck, P_MH, g, (uint64_t)t);
    (mh_step(r, D, x.data(), ll, lp, p.data(), L, P, J, Q) ? na : nr)++;
    if (i % cfg.nskip == 0) record(i / cfg.nskip);
  }
  if (acc) acc[c] = na;
  if (rej) rej[c] = nr;
}

template <class F>
static void parallel_for(int64_t n, int nthreads, F &&f) {
  if (nthreads <= 1 || n <= 1) { for (int64_t i = 0; i < n; ++i) f(i); return; }
  std::vector<std::thread> th;
  int T = (int)std::min<int64_t>(nthreads, n);
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t]() { for (int64_t i = t; i < n; i += T) f(i); });
  for (auto &x : th) x.join();
}

// ===========================================================================
// reversible jump (mcmc.ml:83-153)
// ===========================================================================

struct RjModel {
  LogFn like, prior; Proposal prop;
  int into_kind = 0, nstop = 0; const Tree *tree = nullptr; Proposal into_gauss;
  double p = 0.5; int D = 0;
  explicit RjModel(const mg_rj_model *m) : like(&m->like), prior(&m->prior), prop(&m->prop),
                                            into_kind(m->into.kind), nstop(m->into.nstop),
                                            tree((const Tree *)m->into.tree), p(m->p), D(m->like.dim) {
    if (into_kind == MG_INTO_INDEP_GAUSS) {
      mg_proposal q{MG_PROP_INDEP_GAUSS, D, m->into.params, m->into.nparams};
      into_gauss = Proposal(&q);
    } else if (!tree) throw std::invalid_argument("rj: interp into-proposal without a tree");
  }
  // jintoa / jintob: propose a point of THIS model (test/mcmc_test.ml:175-176)
  bool draw_into(Rng &r, const double *from, double *out) const {
    if (into_kind == MG_INTO_INTERP) return tree->draw(r, nstop, out);
    into_gauss.propose(r, from, out); return true;
  }
  // ljpintoa / ljpintob (test/mcmc_test.ml:177-178: log (Interp.jump_prob ...))
  double log_into(const double *from, const double *to) const {
    if (into_kind == MG_INTO_INTERP) return std::log(tree->jump_prob(to, nstop, nullptr));
    return into_gauss.log_q(from, to);
  }
};

struct RjState { int model; std::vector<double> x; };

static int rj_chain(const CallKey &ck, const RjModel &A, const RjModel &B, const mg_rjmcmc_cfg &cfg,
                    const double *a0, const double *b0, int64_t c, uint8_t *out_model,
                    double *out_samples, int64_t *na_out, int64_t *nb_out, int64_t *acc, double *margin = nullptr,
                    int64_t *cross = nullptr) {
  const int64_t C = cfg.nchains; const int Dm = std::max(A.D, B.D); const int F = Dm + 2;
  tl_margin = HUGE_VAL;
  int64_t ncross = 0, ncross_acc = 0;
  const uint64_t g = cfg.chain_offset + (uint64_t)c;
  const RjModel *M[2] = {&A, &B};
  const double log_p[2] = {std::log(A.p), std::log(B.p)};  // :91
  // mcmc.ml:121-128
  Rng r0(ck, P_RJ_INIT, g, 0);
  bool is_a = r0.uniform() < 0.5;  // F5a
  RjState cur{is_a ? 0 : 1, std::vector<double>(Dm, 0.0)};
  { const double *s = is_a ? a0 : b0; for (int i = 0; i < M[cur.model]->D; ++i) cur.x[i] = s[i]; }
  double ll = M[cur.model]->like(cur.x.data());
  double lp = M[cur.model]->prior(cur.x.data()) + log_p[cur.model];  // :128
  RjState prop{0, std::vector<double>(Dm, 0.0)};
  int64_t na = 0, nb = 0, nacc = 0, t = 0;
  bool bad = false;
  auto step = [&]() {
    Rng r(ck, P_RJ, g, (uint64_t)t);
    // make_mcmc_sampler (:37-56) over the sum type, with the closures of :92-118
    double start_log_post = ll + lp;
    const RjModel &cm = *M[cur.model];
    std::fill(prop.x.begin(), prop.x.end(), 0.0);
    if (r.uniform() < cm.p) {                      // :94,99 stay in model
      prop.model = cur.model; cm.prop.propose(r, cur.x.data(), prop.x.data());
    } else {                                       // :97,102 jump into the other
      prop.model = 1 - cur.model;
      if (!M[prop.model]->draw_into(r, cur.x.data(), prop.x.data())) bad = true;
    }
    const RjModel &pm = *M[prop.model];
    double proposed_like = pm.like(prop.x.data());                       // :113-115
    double proposed_prior = log_p[prop.model] + pm.prior(prop.x.data()); // :116-118
    double proposed_log_posterior = proposed_like + proposed_prior;
    auto ljp = [&](const RjState &x, const RjState &y) {                 // :103-112
      if (x.model == y.model) return log_p[y.model] + M[y.model]->prop.log_q(x.x.data(), y.x.data());
      return log_p[y.model] + M[y.model]->log_into(x.x.data(), y.x.data());
    };
    double log_forward_jump = ljp(cur, prop), log_backward_jump = ljp(prop, cur);
    double log_accept_prob =
        proposed_log_posterior - start_log_post + log_backward_jump - log_forward_jump;
    const double log_u = std::log(r.uniform());
    note_margin(log_u, log_accept_prob);
    if (prop.model != cur.model) ++ncross;
    if (log_u < log_accept_prob) {
      if (prop.model != cur.model) ++ncross_acc;
      cur.model = prop.model; cur.x = prop.x; ll = proposed_like; lp = proposed_prior; ++nacc;
    }
    ++t;
  };
  auto record = [&](int64_t s) {
    (cur.model == 0 ? na : nb)++;
    if (margin) { margin[s * C + c] = tl_margin; tl_margin = HUGE_VAL; }
    if (out_model) out_model[s * C + c] = (uint8_t)cur.model;
    if (out_samples) {
      for (int f = 0; f < Dm; ++f) out_samples[(s * F + f) * C + c] = cur.x[f];
      out_samples[(s * F + Dm) * C + c] = ll; out_samples[(s * F + Dm + 1) * C + c] = lp;
    }
  };
  for (int64_t i = 0; i < cfg.nbin; ++i) step();  // :129-131
  if (cfg.n > 0) record(0);
  for (int64_t i = 1; i <= (cfg.n - 1) * cfg.nskip; ++i) {  // :133-138
    step();
    if (i % cfg.nskip == 0) record(i / cfg.nskip);
  }
  *na_out = na; *nb_out = nb; if (acc) *acc = nacc;
  if (cross) { cross[0] = ncross; cross[1] = ncross_acc; }
  return bad ? 1 : 0;
}

// ---------------------------------------------------------------------------
// k-model reversible jump: an EXTENSION, not a restatement (the reference's sampler is two-model, mcmc.ml:83-87).
// It keeps make_rjmcmc_sampler's structure (mcmc.ml:89-119) over a table of K models and shares its specification
// with the GPU kernel (include/mcmc_gpu.h: mg_rjmcmc_array_k): one uniform walks the model priors cyclically from the
// current model, the last model taking the remainder; initial model uniform over the K; everything else as rj_chain.
// With K = 2 the draws and decisions are rj_chain's.
static int rj_chain_k(const CallKey &ck, const std::vector<RjModel> &Ms, const mg_rjmcmc_cfg &cfg,
                      const double *const *starts, int64_t c, uint8_t *out_model, double *out_samples,
                      int64_t *counts, int64_t *acc, double *margin, int64_t *cross) {
  const int K = (int)Ms.size();
  const int64_t C = cfg.nchains;
  int Dm = 0; for (const auto &m : Ms) Dm = std::max(Dm, m.D);
  const int F = Dm + 2;
  tl_margin = HUGE_VAL;
  int64_t ncross = 0, ncross_acc = 0;
  const uint64_t g = cfg.chain_offset + (uint64_t)c;
  std::vector<double> log_p(K);
  for (int k = 0; k < K; ++k) log_p[k] = std::log(Ms[k].p);
  Rng r0(ck, P_RJ_INIT, g, 0);
  int m0;
  if (K == 2) m0 = r0.uniform() < 0.5 ? 0 : 1;
  else { m0 = (int)(r0.uniform() * (double)K); if (m0 >= K) m0 = K - 1; }
  RjState cur{m0, std::vector<double>(Dm, 0.0)};
  for (int i = 0; i < Ms[m0].D; ++i) cur.x[i] = starts[m0][i];
  double ll = Ms[m0].like(cur.x.data());
  double lp = Ms[m0].prior(cur.x.data()) + log_p[m0];
  RjState prop{0, std::vector<double>(Dm, 0.0)};
  int64_t nacc = 0, t = 0;
  bool bad = false;
  auto pick = [&](int model, double u) {
    if (K == 2) return (u < Ms[model].p) ? model : 1 - model;
    double cum = 0.0; int j = model;
    for (int k = 0; k < K - 1; ++k) {
      cum = cum + Ms[j].p;
      if (u < cum) return j;
      j = (j + 1 == K) ? 0 : j + 1;
    }
    return j;
  };
  auto step = [&]() {
    Rng r(ck, P_RJ, g, (uint64_t)t);
    double start_log_post = ll + lp;
    const RjModel &cm = Ms[cur.model];
    std::fill(prop.x.begin(), prop.x.end(), 0.0);
    prop.model = pick(cur.model, r.uniform());
    if (prop.model == cur.model) cm.prop.propose(r, cur.x.data(), prop.x.data());
    else if (!Ms[prop.model].draw_into(r, cur.x.data(), prop.x.data())) bad = true;
    const RjModel &pm = Ms[prop.model];
    double proposed_like = pm.like(prop.x.data());
    double proposed_prior = log_p[prop.model] + pm.prior(prop.x.data());
    double proposed_log_posterior = proposed_like + proposed_prior;
    auto ljp = [&](const RjState &x, const RjState &y) {
      if (x.model == y.model) return log_p[y.model] + Ms[y.model].prop.log_q(x.x.data(), y.x.data());
      return log_p[y.model] + Ms[y.model].log_into(x.x.data(), y.x.data());
    };
    double log_forward_jump = ljp(cur, prop), log_backward_jump = ljp(prop, cur);
    double log_accept_prob = proposed_log_posterior - start_log_post + log_backward_jump - log_forward_jump;
    const double log_u = std::log(r.uniform());
    note_margin(log_u, log_accept_prob);
    if (prop.model != cur.model) ++ncross;
    if (log_u < log_accept_prob) {
      if (prop.model != cur.model) ++ncross_acc;
      cur.model = prop.model; cur.x = prop.x; ll = proposed_like; lp = proposed_prior; ++nacc;
    }
    ++t;
  };
  auto record = [&](int64_t s) {
    counts[cur.model]++;
    if (margin) { margin[s * C + c] = tl_margin; tl_margin = HUGE_VAL; }
    if (out_model) out_model[s * C + c] = (uint8_t)cur.model;
    if (out_samples) {
      for (int f = 0; f < Dm; ++f) out_samples[(s * F + f) * C + c] = cur.x[f];
      out_samples[(s * F + Dm) * C + c] = ll; out_samples[(s * F + Dm + 1) * C + c] = lp;
    }
  };
  for (int64_t i = 0; i < cfg.nbin; ++i) step();
  if (cfg.n > 0) record(0);
  for (int64_t i = 1; i <= (cfg.n - 1) * cfg.nskip; ++i) {
    step();
    if (i % cfg.nskip == 0) record(i / cfg.nskip);
  }
  if (acc) *acc = nacc;
  if (cross) { cross[0] = ncross; cross[1] = ncross_acc; }
  return bad ? 1 : 0;
}

// ===========================================================================
// evidence.ml
// ===========================================================================

struct Samples {
  int64_t N; int D; const double *pts, *ll, *lp;
};

// evidence.ml:101-107.  out[0]: the reference's left-to-right sum;
// out[1]: the same sum accumulated in long double (80-bit x87 here).
static void harmonic_mean(const double *ll, int64_t n, double out[2]) {
  double linv = 0.0; long double linv_x = 0.0L;
  for (int64_t i = 0; i < n; ++i) {
    double t = 1.0 / std::exp(ll[i]);
    linv = linv + t; linv_x += (long double)t;
  }
  out[0] = (double)n / linv;
  out[1] = (double)((long double)n / linv_x);
}

// evidence.ml:83-89 collect_subvolumes: order = rev(left) @ right.
static void collect_subvolumes(const Tree &t, int32_t id, int nmax, std::vector<int32_t> &out) {
  if (id < 0) return;
  int32_t cnt = t.end[id] - t.begin[id];
  if (cnt < nmax) { out.push_back(id); return; }  // not (length_at_least nmax objs)
  if (t.left[id] < 0) return;                      // both children Empty -> []
  std::vector<int32_t> l, r;
  collect_subvolumes(t, t.left[id], nmax, l);
  collect_subvolumes(t, t.left[id] + 1, nmax, r);
  out.insert(out.end(), l.rbegin(), l.rend());     // List.rev_append left right
  out.insert(out.end(), r.begin(), r.end());
}

static void tight_bounds(const Tree &t, int32_t id, std::vector<double> &lo, std::vector<double> &hi) {
  int32_t b = t.begin[id], e = t.end[id]; int D = t.D;
  lo.assign(t.pt(t.perm[b]), t.pt(t.perm[b]) + D); hi = lo;  // kd_tree.ml:96-110
  for (int32_t k = b + 1; k < e; ++k) {
    const double *c = t.pt(t.perm[k]);
    for (int d = 0; d < D; ++d) { if (c[d] < lo[d]) lo[d] = c[d]; if (c[d] > hi[d]) hi[d] = c[d]; }
  }
}

// Tree over a subset of samples (rows `idx` in this order), tight root box.
static void tree_of_subset(const Samples &s, const std::vector<int64_t> &idx, int min_split, Tree &t) {
  t.D = s.D; t.N = (int64_t)idx.size(); t.min_split = min_split;
  t.pts.resize((size_t)t.N * s.D);
  for (int64_t i = 0; i < t.N; ++i) std::memcpy(&t.pts[(size_t)i * s.D], s.pts + idx[i] * s.D, sizeof(double) * s.D);
  t.low.assign(s.D, 0.0); t.high.assign(s.D, 0.0);
  if (t.N > 0) {  // bounds_of_objects, evidence.ml:164,206
    for (int d = 0; d < s.D; ++d) t.low[d] = t.high[d] = t.pts[d];
    for (int64_t i = 1; i < t.N; ++i)
      for (int d = 0; d < s.D; ++d) {
        double c = t.pts[(size_t)i * s.D + d];
        if (c < t.low[d]) t.low[d] = c; if (c > t.high[d]) t.high[d] = c;
      }
  }
  t.build();
}

// evidence.ml:148-165 evidence_direct.  full_tree = 0 stops splitting below n
// objects (same cells: collect_subvolumes never descends past them).
static int evidence_direct(const Samples &s, int n, int full_tree, double out[2], int64_t *ncells) {
  if (s.N <= 0) return fail(MG_EINVAL, "bounds_of_objects: no objects");
  // array_to_list_remove_dups :143-146: List.sort (stable) by coordinates,
  // then rev_remove_dups :126-141 (keeps the last of each run, reversed).
  std::vector<int64_t> order(s.N);
  for (int64_t i = 0; i < s.N; ++i) order[i] = i;
  auto cmp = [&](int64_t a, int64_t b) {
    const double *x = s.pts + a * s.D, *y = s.pts + b * s.D;
    for (int d = 0; d < s.D; ++d) { int c = fcompare(x[d], y[d]); if (c) return c < 0; }
    return false;
  };
  auto eq = [&](int64_t a, int64_t b) { return !cmp(a, b) && !cmp(b, a); };
  std::stable_sort(order.begin(), order.end(), cmp);
  std::vector<int64_t> kept;
  for (int64_t i = 0; i < s.N; ++i)
    if (i == s.N - 1 || !eq(order[i], order[i + 1])) kept.push_back(order[i]);
  std::reverse(kept.begin(), kept.end());
  Tree t; tree_of_subset(s, kept, full_tree ? 2 : n, t);
  std::vector<int32_t> cells; collect_subvolumes(t, 0, n, cells);
  if (ncells) *ncells = (int64_t)cells.size();
  double integral = 0.0; long double integral_x = 0.0L;
  std::vector<double> lo, hi;
  for (int32_t id : cells) {  // :150-160
    tight_bounds(t, id, lo, hi);
    double vol = Tree::bounds_volume(lo.data(), hi.data(), s.D);
    double sum = 0.0; int cnt = 0;                 // mean_sample posterior :122-124
    for (int32_t k = t.begin[id]; k < t.end[id]; ++k) {
      int64_t src = kept[t.perm[k]];
      sum = sum + std::exp(s.ll[src] + s.lp[src]); ++cnt;
    }
    double post = sum / (double)cnt;
    integral = integral + vol * post; integral_x += (long double)(vol * post);
  }
  out[0] = integral; out[1] = (double)integral_x;
  return MG_OK;
}

// evidence.ml:202-221 evidence_lebesgue.
static int evidence_lebesgue(const Samples &s, int n, double eps, int full_tree, double out[2],
                             int64_t *nkept_out, int64_t *ncells) {
  if (s.N <= 0) return fail(MG_EINVAL, "bounds_of_objects: no objects");
  // collect_samples_up_to_eps :167-180: List.fast_sort (stable) by -ll
  std::vector<int64_t> order(s.N);
  for (int64_t i = 0; i < s.N; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(),
                   [&](int64_t a, int64_t b) { return fcompare(-s.ll[a], -s.ll[b]) < 0; });
  int64_t m = s.N;
  for (int64_t i = 0; i + 1 < s.N; ++i) {
    double ilx = std::exp(-s.ll[order[i]]), ily = std::exp(-s.ll[order[i + 1]]);
    double delta = ily - ilx;
    if (!(delta >= 0.0)) return fail(MG_EFAIL, "Assert_failure evidence.ml:175");
    if (delta > eps) { m = i + 1; break; }
  }
  // mean_inv_like :182-189
  double tot_il = 0.0; long double tot_x = 0.0L;
  for (int64_t i = 0; i < m; ++i) { double t = std::exp(-s.ll[order[i]]); tot_il = tot_il + t; tot_x += t; }
  double mean_il = tot_il / (double)m; long double mean_x = tot_x / (long double)m;
  // remove_dups_rev :191-200 (keeps the last of each equal-ll run, reversed)
  std::vector<int64_t> kept;
  for (int64_t i = 0; i < m; ++i)
    if (i == m - 1 || !(s.ll[order[i]] == s.ll[order[i + 1]])) kept.push_back(order[i]);
  std::reverse(kept.begin(), kept.end());
  if (nkept_out) *nkept_out = (int64_t)kept.size();
  Tree t; tree_of_subset(s, kept, full_tree ? 2 : n, t);
  std::vector<int32_t> cells; collect_subvolumes(t, 0, n, cells);
  if (ncells) *ncells = (int64_t)cells.size();
  double pm = 0.0; long double pm_x = 0.0L;
  std::vector<double> lo, hi, lps;
  for (int32_t id : cells) {  // :209-220
    tight_bounds(t, id, lo, hi);
    double vol = Tree::bounds_volume(lo.data(), hi.data(), s.D);
    lps.clear();
    for (int32_t k = t.begin[id]; k < t.end[id]; ++k) lps.push_back(s.lp[kept[t.perm[k]]]);
    std::stable_sort(lps.begin(), lps.end());      // median_sample :109-120
    size_t c = lps.size();
    double med = (c % 2 == 0) ? 0.5 * (lps[c / 2 - 1] + lps[c / 2]) : lps[c / 2];
    double prior = std::exp(med);
    pm = pm + prior * vol; pm_x += (long double)(prior * vol);
  }
  out[0] = pm / mean_il; out[1] = (double)(pm_x / mean_x);
  return MG_OK;
}

// ===========================================================================
// nested.ml
// ===========================================================================

struct LivePt { std::vector<double> x; double ll, lp; };

// Shrinkage schedule.  The reference retires one point per iteration with
// factor (1 - 1/nlive) (nested.ml:84,96,131).  Retiring the K lowest at once,
// the j-th of a batch leaves nlive-j points above it: factor 1 - 1/(nlive-j).
// log X_i (volume before retiring point i = b*K + j) = b*S_K + s_j.
struct Shrink {
  int nlive, K; std::vector<double> s; double S;
  Shrink(int nlive_, int K_) : nlive(nlive_), K(K_), s(K_ + 1, 0.0) {
    for (int j = 0; j < K; ++j) s[j + 1] = s[j] + std::log1p(-(1.0 / (double)(nlive - j)));
    S = s[K];
  }
  double log_x(int64_t i) const { return (double)(i / K) * S + s[i % K]; }
  double vol_fraction(int64_t i) const { return 1.0 / (double)(nlive - (int)(i % K)); }
};

// nested.ml:81-120 evidence_error_and_weights (K = 1: operation for operation).
static void nested_weights(const double *ll, int64_t n, int nlive, int K, double *log_ev,
                           double *log_dev, double *wts) {
  Shrink sh(nlive, K);
  const double log_half = -0.69314718055994530942;
  for (int64_t i = 0; i < n; ++i) wts[i] = NEG_INF;
  double low = NEG_INF, high = NEG_INF;
  int64_t ilive = n - nlive;
  for (int64_t i = 0; i < ilive; ++i) {
    double log_dv = std::log(sh.vol_fraction(i)) + sh.log_x(i);
    double log_dlow = log_dv + ll[i], log_dhigh = log_dv + ll[i + 1];
    low = log_sum_logs(low, log_dlow); high = log_sum_logs(high, log_dhigh);
    wts[i] = log_sum_logs(wts[i], log_half + log_dlow);
    wts[i + 1] = log_sum_logs(wts[i + 1], log_half + log_dhigh);
  }
  double log_dv = std::log(1.0 / (double)nlive) + sh.log_x(ilive - 1);  // :97
  for (int64_t i = ilive; i < n; ++i) {
    if (i == 0) continue;
    double log_dlow = log_dv + ll[i - 1], log_dhigh = log_dv + ll[i];
    low = log_sum_logs(low, log_dlow); high = log_sum_logs(high, log_dhigh);
    wts[i - 1] = log_sum_logs(wts[i - 1], log_half + log_dlow);
    wts[i] = log_sum_logs(wts[i], log_half + log_dhigh);
  }
  *log_ev = log_half + log_sum_logs(low, high);
  *log_dev = high + std::log1p(-std::exp(low - high));
  for (int64_t i = 0; i < n; ++i) wts[i] = wts[i] - *log_ev;
}

// mcmc.ml:198-218 differential_evolution_proposal over the live set.
static void de_propose(Rng &r, const std::vector<LivePt> &live, double mode_hop, int D,
                       const double *z, double *out) {
  uint64_t n = live.size();
  uint64_t i = r.below(n), j;
  do { j = r.below(n); } while (j == i);
  double d;
  if (mode_hop != 0.0 && r.uniform() < mode_hop) d = 1.0;
  else { double sigma = 2.38 / std::sqrt(2.0 * (double)D); d = draw_gaussian(r, 0.0, sigma); }
  const double *x = live[i].x.data(), *y = live[j].x.data();
  for (int k = 0; k < D; ++k) out[k] = z[k] + d * (y[k] - x[k]);
}

// nested.ml:122-146 nested_evidence with K-at-a-time replacement (K = 1 is
// the reference's schedule).
static int nested_evidence(uint64_t seed, uint64_t epoch, const LogFn &like, const LogFn &prior,
                           const double *plo, const double *phi, const mg_nested_cfg &cfg,
                           double *log_ev, double *log_dev, int64_t *npts, double *pts, double *llo,
                           double *lpo, double *logw) {
  const int D = cfg.dim, nlive = cfg.nlive, K = cfg.batch;
  if (K < 1 || K >= nlive) return fail(MG_EINVAL, "nested: need 1 <= batch < nlive");
  CallKey ck = derive_key(seed, epoch);
  std::vector<LivePt> live(nlive);
  for (int i = 0; i < nlive; ++i) {  // :126-130, draw_prior = per-dim draw_uniform
    Rng r(ck, P_NEST_INIT, (uint64_t)i, 0);
    live[i].x.resize(D);
    for (int d = 0; d < D; ++d) live[i].x[d] = draw_uniform(r, plo[d], phi[d]);
    live[i].ll = like(live[i].x.data()); live[i].lp = prior(live[i].x.data());
  }
  auto by_ll = [](const LivePt &a, const LivePt &b) { return fcompare(a.ll, b.ll) < 0; };
  std::stable_sort(live.begin(), live.end(), by_ll);  // :132
  std::vector<LivePt> retired;
  double log_vol = 0.0, log_int = NEG_INF;
  int64_t R = 0;
  std::vector<double> p(D);
  for (;;) {
    double thr = live[K - 1].ll;
    std::vector<LivePt> fresh(K);
    for (int j = 0; j < K; ++j) {  // draw_new_live_point :50-74
      uint64_t rid = (uint64_t)(R + j);
      Rng rs(ck, P_NEST_START, rid, 0);
      const LivePt &st = live[(K - 1) + (int)rs.below((uint64_t)(nlive - K + 1))];  // :63
      std::vector<double> x = st.x;
      auto mcmc_logl = [&](const double *pt) { double l = like(pt); return l >= thr ? prior(pt) : NEG_INF; };
      double cl = mcmc_logl(x.data()), cp = 0.0;
      for (int s = 0; s < cfg.nmcmc; ++s) {
        Rng r(ck, P_NEST_MCMC, rid, (uint64_t)s);
        mh_step(r, D, x.data(), cl, cp, p.data(), mcmc_logl, [](const double *) { return 0.0; },
                [&](Rng &rr, const double *a, double *b) { de_propose(rr, live, cfg.mode_hopping_frac, D, a, b); },
                [](const double *, const double *) { return 0.0; });
      }
      fresh[j].x = x; fresh[j].ll = like(x.data()); fresh[j].lp = prior(x.data());
      if (!(fresh[j].ll >= thr)) return fail(MG_EFAIL, "Error in draw_new_live_point: new log(L) below threshold");
    }
    for (int j = 0; j < K; ++j) {  // nested_loop :133-145 for each retired point
      double m = (double)(nlive - j);
      double vol_fraction = 1.0 / m;
      double log_new_vol = log_vol + std::log1p(-vol_fraction);
      double log_dv = log_vol + vol_fraction;      // :140, quirk F5d (not log vol_fraction)
      log_int = log_sum_logs(log_int, live[j].ll + log_dv);
      log_vol = log_new_vol;
      retired.push_back(live[j]);
      live[j] = fresh[j];
    }
    R += K;
    std::stable_sort(live.begin(), live.end(), by_ll);  // replace_live_point :26-43
    // remaining_integral_negligable :45-48
    double log_live_estimate = log_vol + live[nlive - 1].ll;
    if (log_live_estimate - log_sum_logs(log_int, log_live_estimate) <= std::log(cfg.epsrel)) break;
    if ((int64_t)retired.size() + nlive + K > cfg.max_points) return fail(MG_EFAIL, "nested: max_points too small");
  }
  int64_t n = (int64_t)retired.size() + nlive;
  if (n > cfg.max_points) return fail(MG_EFAIL, "nested: max_points too small");
  for (int64_t i = 0; i < n; ++i) {
    const LivePt &q = i < (int64_t)retired.size() ? retired[i] : live[i - retired.size()];
    if (pts) std::memcpy(pts + i * D, q.x.data(), sizeof(double) * D);
    llo[i] = q.ll; if (lpo) lpo[i] = q.lp;
  }
  *npts = n;
  nested_weights(llo, n, nlive, K, log_ev, log_dev, logw);
  return MG_OK;
}

// ===========================================================================
// C API (ctypes)
// ===========================================================================
#define OG_TRY try {
#define OG_CATCH                                                             \
  }                                                                          \
  catch (const std::invalid_argument &e) { return fail(MG_EINVAL, e.what()); } \
  catch (const std::exception &e) { return fail(MG_EFAIL, e.what()); }

extern "C" {

const char *og_last_error() { return g_err.c_str(); }

void og_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { Philox::block(ctr, key, out); }

// draws j = 0..n-1 of (purpose, g, step): uniforms and raw lanes
void og_rng_stream(uint64_t seed, uint64_t epoch, uint32_t purpose, uint64_t g, uint64_t step, int n,
                   double *u, uint64_t *lanes) {
  CallKey ck = derive_key(seed, epoch);
  if (u) { Rng r(ck, purpose, g, step); for (int i = 0; i < n; ++i) u[i] = r.uniform(); }
  if (lanes) { Rng r(ck, purpose, g, step); for (int i = 0; i < n; ++i) lanes[i] = r.lane(); }
}

int og_logfn_eval(const mg_logfn *fn, const double *x, int64_t M, double *out) {
  OG_TRY
  LogFn f(fn);
  for (int64_t i = 0; i < M; ++i) out[i] = f(x + i * fn->dim);
  return MG_OK;
  OG_CATCH
}

// margin: [n][C] or null -- min |log u - log_accept_prob| over the steps leading to each slot (test diagnostic)
int og_mcmc_array_m(uint64_t seed, uint64_t epoch, const mg_logfn *like, const mg_logfn *prior,
                    const mg_proposal *prop, const mg_mcmc_cfg *cfg, const double *x0, double *out,
                    int64_t *acc, int64_t *rej, int nthreads, double *margin) {
  OG_TRY
  LogFn L(like), P(prior); Proposal J(prop);
  if (cfg->nskip < 1 || cfg->n < 0 || cfg->nbin < 0) return fail(MG_EINVAL, "mcmc_array: bad nbin/nskip/n");
  CallKey ck = derive_key(seed, epoch);
  parallel_for(cfg->nchains, nthreads, [&](int64_t c) {
    const double *s = cfg->x0_shared ? x0 : x0 + c * cfg->dim;
    mcmc_chain(ck, L, P, J, *cfg, s, c, out, acc, rej, margin);
  });
  return MG_OK;
  OG_CATCH
}
int og_mcmc_array(uint64_t seed, uint64_t epoch, const mg_logfn *like, const mg_logfn *prior,
                  const mg_proposal *prop, const mg_mcmc_cfg *cfg, const double *x0, double *out,
                  int64_t *acc, int64_t *rej, int nthreads) {
  return og_mcmc_array_m(seed, epoch, like, prior, prop, cfg, x0, out, acc, rej, nthreads, nullptr);
}

// margin: [n][C] or null (as og_mcmc_array_m); out_cross: {cross-model proposals, accepted} or null
int og_rjmcmc_array_m(uint64_t seed, uint64_t epoch, const mg_rj_model *A, const mg_rj_model *B,
                      const mg_rjmcmc_cfg *cfg, const double *a0, const double *b0, uint8_t *out_model,
                      double *out_samples, int64_t out_counts[2], int64_t *out_accept, int nthreads,
                      double *margin, int64_t *out_cross) {
  OG_TRY
  // mcmc.ml:90 assert(pa +. pb -. 1.0 < sqrt epsilon_float)  (one-sided)
  if (!(A->p + B->p - 1.0 < std::sqrt(2.220446049250313e-16))) return fail(MG_EFAIL, "Assert_failure mcmc.ml:90");
  RjModel a(A), b(B);
  CallKey ck = derive_key(seed, epoch);
  std::vector<int64_t> na(cfg->nchains), nb(cfg->nchains), ac(cfg->nchains);
  std::vector<int> bad(cfg->nchains, 0);
  std::vector<int64_t> cr((size_t)cfg->nchains * 2, 0);
  parallel_for(cfg->nchains, nthreads, [&](int64_t c) {
    bad[c] = rj_chain(ck, a, b, *cfg, a0, b0, c, out_model, out_samples, &na[c], &nb[c], &ac[c], margin, &cr[2 * c]);
  });
  int64_t ta = 0, tb = 0, tacc = 0, c0 = 0, c1 = 0;
  for (int64_t c = 0; c < cfg->nchains; ++c) {
    ta += na[c]; tb += nb[c]; tacc += ac[c]; c0 += cr[2 * c]; c1 += cr[2 * c + 1];
    if (bad[c]) return fail(MG_EFAIL, "draw: empty tree");
  }
  out_counts[0] = ta; out_counts[1] = tb;
  if (out_accept) *out_accept = tacc;
  if (out_cross) { out_cross[0] = c0; out_cross[1] = c1; }
  return MG_OK;
  OG_CATCH
}
int og_rjmcmc_array(uint64_t seed, uint64_t epoch, const mg_rj_model *A, const mg_rj_model *B,
                    const mg_rjmcmc_cfg *cfg, const double *a0, const double *b0, uint8_t *out_model,
                    double *out_samples, int64_t out_counts[2], int64_t *out_accept, int nthreads) {
  return og_rjmcmc_array_m(seed, epoch, A, B, cfg, a0, b0, out_model, out_samples, out_counts, out_accept, nthreads,
                           nullptr, nullptr);
}

// k-model extension (see rj_chain_k).  out_counts: [nmodels]; margin: [n][C] or null; out_cross: {proposed, accepted}
int og_rjmcmc_array_k(uint64_t seed, uint64_t epoch, const mg_rj_model *models, int32_t nmodels,
                      const mg_rjmcmc_cfg *cfg, const double *const *starts, uint8_t *out_model, double *out_samples,
                      int64_t *out_counts, int64_t *out_accept, int nthreads, double *margin, int64_t *out_cross) {
  OG_TRY
  if (nmodels < 2 || nmodels > MG_RJ_MAX_MODELS) return fail(MG_EINVAL, "rjmcmc_array_k: 2..MG_RJ_MAX_MODELS models");
  double psum = 0.0;
  std::vector<RjModel> Ms;
  for (int k = 0; k < nmodels; ++k) { psum = psum + models[k].p; Ms.emplace_back(models + k); }
  if (!(psum - 1.0 < std::sqrt(2.220446049250313e-16))) return fail(MG_EFAIL, "Assert_failure mcmc.ml:90");
  CallKey ck = derive_key(seed, epoch);
  const int64_t C = cfg->nchains;
  std::vector<int64_t> cnt((size_t)C * nmodels, 0), ac(C, 0), cr((size_t)C * 2, 0);
  std::vector<int> bad(C, 0);
  parallel_for(C, nthreads, [&](int64_t c) {
    bad[c] = rj_chain_k(ck, Ms, *cfg, starts, c, out_model, out_samples, &cnt[(size_t)c * nmodels], &ac[c], margin, &cr[2 * c]);
  });
  for (int k = 0; k < nmodels; ++k) out_counts[k] = 0;
  int64_t tacc = 0, c0 = 0, c1 = 0;
  for (int64_t c = 0; c < C; ++c) {
    for (int k = 0; k < nmodels; ++k) out_counts[k] += cnt[(size_t)c * nmodels + k];
    tacc += ac[c]; c0 += cr[2 * c]; c1 += cr[2 * c + 1];
    if (bad[c]) return fail(MG_EFAIL, "draw: empty tree");
  }
  if (out_accept) *out_accept = tacc;
  if (out_cross) { out_cross[0] = c0; out_cross[1] = c1; }
  return MG_OK;
  OG_CATCH
}

// ---- kd-tree / interpolate ------------------------------------------------
int og_kdtree_build(const double *pts, int64_t N, int32_t D, const double *low, const double *high,
                    int32_t min_split, void **out) {
  OG_TRY
  if (N <= 0 || D <= 0) return fail(MG_EINVAL, "tree_of_objects: no objects");
  for (int64_t i = 0; i < N * D; ++i) if (pts[i] != pts[i]) return fail(MG_EINVAL, "NaN coordinate");
  Tree *t = new Tree;
  t->D = D; t->N = N; t->min_split = min_split < 2 ? 2 : min_split;
  t->pts.assign(pts, pts + N * D);
  t->low.assign(low, low + D); t->high.assign(high, high + D);
  t->build();
  *out = t;
  return MG_OK;
  OG_CATCH
}
void og_kdtree_destroy(void *t) { delete (Tree *)t; }
int og_kdtree_info(const void *tp, int64_t *npoints, int32_t *dim, int64_t *nnodes, int32_t *nlevels) {
  const Tree *t = (const Tree *)tp;
  if (npoints) *npoints = t->N; if (dim) *dim = t->D;
  if (nnodes) *nnodes = t->nnodes(); if (nlevels) *nlevels = t->nlevels;
  return MG_OK;
}
int og_kdtree_export(const void *tp, int32_t *split_dim, double *split_val, int32_t *left, int32_t *begin,
                     int32_t *end, int32_t *perm) {
  const Tree *t = (const Tree *)tp; size_t n = (size_t)t->nnodes();
  if (split_dim) std::memcpy(split_dim, t->dim.data(), n * 4);
  if (split_val) std::memcpy(split_val, t->split.data(), n * 8);
  if (left) std::memcpy(left, t->left.data(), n * 4);
  if (begin) std::memcpy(begin, t->begin.data(), n * 4);
  if (end) std::memcpy(end, t->end.data(), n * 4);
  if (perm) std::memcpy(perm, t->perm.data(), (size_t)t->N * 4);
  return MG_OK;
}
double og_bounds_volume(const double *lo, const double *hi, int32_t D) { return Tree::bounds_volume(lo, hi, D); }
int og_interp_find_cell(const void *tp, const double *q, int64_t M, int32_t nstop, int32_t *out_node,
                        double *out_lo, double *out_hi) {
  const Tree *t = (const Tree *)tp; std::vector<double> lo(t->D), hi(t->D);
  for (int64_t i = 0; i < M; ++i) {
    out_node[i] = t->find_cell(q + i * t->D, nstop, lo.data(), hi.data());
    if (out_lo) std::memcpy(out_lo + i * t->D, lo.data(), 8 * t->D);
    if (out_hi) std::memcpy(out_hi + i * t->D, hi.data(), 8 * t->D);
  }
  return MG_OK;
}
int og_interp_jump_prob(const void *tp, const double *q, int64_t M, int32_t nstop, double *out_prob) {
  const Tree *t = (const Tree *)tp;
  for (int64_t i = 0; i < M; ++i) out_prob[i] = t->jump_prob(q + i * t->D, nstop, nullptr);
  return MG_OK;
}
int og_interp_draw(uint64_t seed, uint64_t epoch, const void *tp, int64_t M, int32_t nstop, double *out) {
  const Tree *t = (const Tree *)tp; CallKey ck = derive_key(seed, epoch);
  for (int64_t i = 0; i < M; ++i) {
    Rng r(ck, P_DRAW, (uint64_t)i, 0);
    if (!t->draw(r, nstop, out + i * t->D)) return fail(MG_EFAIL, "draw_high_level: encountered empty tree!");
  }
  return MG_OK;
}

// ---- evidence / stats -----------------------------------------------------
int og_evidence_harmonic_mean(const double *ll, int64_t N, double out[2]) { harmonic_mean(ll, N, out); return MG_OK; }
int og_evidence_lebesgue(const double *pts, const double *ll, const double *lp, int64_t N, int32_t D, int32_t n,
                         double eps, int32_t full_tree, double out[2], int64_t *nkept, int64_t *ncells) {
  OG_TRY
  Samples s{N, D, pts, ll, lp};
  return evidence_lebesgue(s, n, eps, full_tree, out, nkept, ncells);
  OG_CATCH
}
int og_evidence_direct(const double *pts, const double *ll, const double *lp, int64_t N, int32_t D, int32_t n,
                       int32_t full_tree, double out[2], int64_t *ncells) {
  OG_TRY
  Samples s{N, D, pts, ll, lp};
  return evidence_direct(s, n, full_tree, out, ncells);
  OG_CATCH
}
// bin/harmonic_evidence.ml:41-52: bootstrap replicates of the harmonic-mean evidence
// (resample n indices with Random.int n, recompute).  evs is returned UNSORTED.
int og_harmonic_bootstrap(uint64_t seed, uint64_t epoch, const double *ll, int64_t n, int32_t nbstrap, double *evs) {
  CallKey ck = derive_key(seed, epoch);
  for (int32_t b = 0; b < nbstrap; ++b) {
    Rng r(ck, P_BOOT, (uint64_t)b, 0);
    double linv = 0.0;
    for (int64_t i = 0; i < n; ++i) linv = linv + 1.0 / std::exp(ll[r.below((uint64_t)n)]);
    evs[b] = (double)n / linv;
  }
  return MG_OK;
}
// mcmc.ml:74-81 remove_repeat_samples on rows [n][F] comparing the first D
int64_t og_remove_repeat_samples(const double *rows, int64_t n, int32_t D, double *out) {
  int F = D + 2; int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    bool keep = (i == 0);
    if (!keep) for (int d = 0; d < D; ++d) if (!(rows[i * F + d] == rows[(i - 1) * F + d])) { keep = true; break; }
    if (keep) { std::memcpy(out + k * F, rows + i * F, 8 * F); ++k; }
  }
  return k;
}
double og_stats_mean(const double *x, int64_t n) { return mean(x, n); }
double og_stats_std(const double *x, int64_t n, int have_mean, double mu) { return stdev(x, n, have_mean != 0, mu); }
void og_stats_multi_mean(const double *xs, int64_t n, int32_t D, double *out) { multi_mean(xs, n, D, out); }
void og_stats_multi_std(const double *xs, int64_t n, int32_t D, const double *mean_or_null, double *out) {
  multi_std(xs, n, D, mean_or_null, out);
}
void og_stats_autocorrelation(const double *x, int64_t n, int32_t nslides, double *r, double *length) {
  slow_autocorrelation(x, n, nslides, r);
  if (length) { double L = 1.0; for (int i = 1; i < nslides && r[i] > 0.0; ++i) L += 2.0 * r[i]; *length = L; }
}
double og_log_sum_logs(double a, double b) { return log_sum_logs(a, b); }
double og_log_gaussian(double mu, double sigma, double x) { return log_gaussian(mu, sigma, x); }
double og_log_cauchy(double x0, double g, double x) { return log_cauchy(x0, g, x); }
double og_log_lognormal(double mu, double sigma, double x) { return log_lognormal(mu, sigma, x); }
void og_draw_gaussian(uint64_t seed, uint64_t epoch, double mu, double sigma, int64_t n, double *out) {
  CallKey ck = derive_key(seed, epoch);
  for (int64_t i = 0; i < n; ++i) { Rng r(ck, P_DRAW, (uint64_t)i, 1); out[i] = draw_gaussian(r, mu, sigma); }
}
// stats.ml:89-91,113-128 in bulk, kinds as include/mcmc_gpu.h MG_DRAW_*: 0 uniform a b, 1 gaussian mu sigma, 2 cauchy x0 gamma
void og_stats_draw(uint64_t seed, uint64_t epoch, int32_t kind, double a, double b, int64_t n, double *out) {
  CallKey ck = derive_key(seed, epoch);
  for (int64_t i = 0; i < n; ++i) {
    Rng r(ck, P_DRAW, (uint64_t)i, 1);
    out[i] = kind == 0 ? draw_uniform(r, a, b) : kind == 1 ? draw_gaussian(r, a, b) : draw_cauchy(r, a, b);
  }
}
// mcmc.ml:198-218 on a fixed sample table [n][D]; M proposals from `z`
void og_de_proposals(uint64_t seed, uint64_t epoch, const double *table, int64_t n, int32_t D, double mode_hop,
                     const double *z, int64_t M, double *out) {
  CallKey ck = derive_key(seed, epoch);
  std::vector<LivePt> live(n);
  for (int64_t i = 0; i < n; ++i) live[i].x.assign(table + i * D, table + (i + 1) * D);
  for (int64_t m = 0; m < M; ++m) { Rng r(ck, P_NEST_MCMC, (uint64_t)m, 0); de_propose(r, live, mode_hop, D, z, out + m * D); }
}

// ---- nested ---------------------------------------------------------------
int og_nested_evidence(uint64_t seed, uint64_t epoch, const mg_logfn *like, const mg_logfn *prior,
                       const double *plo, const double *phi, const mg_nested_cfg *cfg, double *log_ev,
                       double *log_dev, int64_t *npts, double *pts, double *ll, double *lp, double *logw) {
  OG_TRY
  LogFn L(like), P(prior);
  return nested_evidence(seed, epoch, L, P, plo, phi, *cfg, log_ev, log_dev, npts, pts, ll, lp, logw);
  OG_CATCH
}
int og_nested_weights(const double *ll, int64_t n, int32_t nlive, int32_t batch, double *log_ev, double *log_dev,
                      double *logw) {
  if (n < nlive || batch < 1 || batch >= nlive) return fail(MG_EINVAL, "nested_weights: bad sizes");
  nested_weights(ll, n, nlive, batch, log_ev, log_dev, logw);
  return MG_OK;
}
// nested.ml:148-150
double og_nested_log_total_error(double log_ev, double log_dev, int32_t nlive) {
  double log_rel_error2 = -std::log((double)nlive);
  return 0.5 * log_sum_logs(2.0 * log_dev, log_rel_error2 + 2.0 * log_ev);
}
// nested.ml:152-178 posterior_samples: indices of n draws
void og_nested_posterior_indices(uint64_t seed, uint64_t epoch, const double *logw, int64_t npts, int64_t n,
                                 int64_t *out_idx) {
  CallKey ck = derive_key(seed, epoch);
  std::vector<double> sw(npts);
  sw[0] = std::exp(logw[0]);
  for (int64_t i = 1; i < npts; ++i) sw[i] = std::exp(logw[i]) + sw[i - 1];
  for (int64_t k = 0; k < n; ++k) {
    Rng r(ck, P_POST, (uint64_t)k, 0);
    double x = r.uniform();
    int64_t idx;
    if (x <= sw[0]) idx = 0;
    else { int64_t lo = 0, hi = npts - 1; while (hi - lo > 1) { int64_t mid = (lo + hi) / 2; if (x <= sw[mid]) hi = mid; else lo = mid; } idx = hi; }
    out_idx[k] = idx;
  }
}

}  // extern "C"
