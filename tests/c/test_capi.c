/* test_capi.c -- a plain C program that drives libmcmcgpu.so exactly as ocaml/mcmc_gpu_stubs.c does (same structs,
 * same call sequences, host buffers laid out as the Bigarrays are), without Python or ctypes in between:
 *   mcmcgpu_mcmc_array_native   -> mg_mcmc_array (chain-major layout, shared start point)
 *   mcmcgpu_interp_make / _jump_prob / _draw
 *   mcmcgpu_rjmcmc_array_native -> mg_rjmcmc_array (Interp jumps into both models)
 *   mcmcgpu_evidence_*          -> mg_evidence_harmonic_mean / _lebesgue / _direct
 *   mcmcgpu_nested_evidence_native -> mg_nested_evidence, mg_nested_log_total_error, mg_nested_posterior_indices
 *   mcmcgpu_stats_*             -> mg_stats_multi_mean / _multi_std / mg_stats_draw
 *   mcmcgpu_rjmcmc_array_k_native -> mg_rjmcmc_array_k;  mcmcgpu_enclosing_ellipse -> mg_ellipse_*
 * and checks the reference's known answers (test/mcmc_test.ml:150-182 ratio 4, test/nested_test.ml:23-39 evidence 1).
 * Built by __graft_entry__.build() with gcc; run on the GPU box by tests/test_abi.py::test_c_program. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcmc_gpu.h"

#define CHECK(ctx, call)                                                                            \
  do { int rc_ = (call); if (rc_ != MG_OK) { fprintf(stderr, "FAIL %s: %s\n", #call, mg_last_error(ctx)); return 1; } } while (0)
#define EXPECT(cond) do { if (!(cond)) { fprintf(stderr, "FAIL expectation: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

int main(void) {
  mg_ctx *ctx = NULL;
  if (mg_ctx_create(0, 20111104ull, &ctx) != MG_OK) { fprintf(stderr, "FAIL no GPU context\n"); return 1; }
  EXPECT(mg_abi_version() == MG_ABI_VERSION);

  /* ---- Mcmc.mcmc_array: unit square and the central 0.5 x 0.5 square (test/mcmc_test.ml:150-169) ---------------- */
  const int D = 2;
  double box1[5] = {0, 0, 1, 1, 0.0}, box2[5] = {0.25, 0.25, 0.75, 0.75, 0.0};
  double wrap[6] = {0, 0, 1, 1, 0.5, 0.5};
  mg_logfn prior = {MG_FN_BOX_CLOSED, D, 1.0, box1, 5}, like1 = {MG_FN_BOX_CLOSED, D, 1.0, box1, 5},
           like2 = {MG_FN_BOX_CLOSED, D, 1.0, box2, 5};
  mg_proposal prop = {MG_PROP_WRAP, D, wrap, 6};
  const int64_t n = 5000, C = 4;
  mg_mcmc_cfg cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.nchains = C; cfg.dim = D; cfg.layout = MG_LAYOUT_CHAIN_MAJOR; cfg.nskip = 20; cfg.n = n; cfg.x0_shared = 1;
  double x0[2] = {0.5, 0.5};
  double *s1 = malloc(sizeof(double) * C * n * (D + 2)), *s2 = malloc(sizeof(double) * C * n * (D + 2));
  CHECK(ctx, mg_mcmc_array(ctx, &like1, &prior, &prop, &cfg, x0, s1, NULL, NULL));
  CHECK(ctx, mg_mcmc_array(ctx, &like2, &prior, &prop, &cfg, x0, s2, NULL, NULL));
  int64_t na = 0, nr = 0;
  CHECK(ctx, mg_get_counters(ctx, &na, &nr));
  EXPECT(na + nr == 2 * C * (n - 1) * 20 && na > 0 && nr > 0);
  /* pooled points [C*n][D] for Interp.make */
  double *p1 = malloc(sizeof(double) * C * n * D), *p2 = malloc(sizeof(double) * C * n * D);
  for (int64_t i = 0; i < C * n; ++i)
    for (int d = 0; d < D; ++d) { p1[i * D + d] = s1[i * (D + 2) + d]; p2[i * D + d] = s2[i * (D + 2) + d]; }
  for (int64_t i = 0; i < C * n; ++i) EXPECT(p2[i * D] >= 0.25 && p2[i * D] <= 0.75 && p2[i * D + 1] >= 0.25 && p2[i * D + 1] <= 0.75);

  /* ---- Interpolate_pdf.make / jump_prob / draw ------------------------------------------------------------------ */
  double lo[2] = {0, 0}, hi[2] = {1, 1};
  mg_kdtree *t1 = NULL, *t2 = NULL;
  CHECK(ctx, mg_kdtree_build(ctx, p1, C * n, D, lo, hi, 2, &t1));
  CHECK(ctx, mg_kdtree_build(ctx, p2, C * n, D, lo, hi, 2, &t2));
  double q[4] = {0.5, 0.5, 0.1, 0.9}, jp[2], dr[6];
  CHECK(ctx, mg_interp_jump_prob(ctx, t1, q, 2, 0, jp));
  EXPECT(jp[0] > 0.1 && jp[0] < 10.0 && jp[1] > 0.1 && jp[1] < 10.0);       /* density of a uniform unit square ~ 1 */
  CHECK(ctx, mg_interp_draw(ctx, t2, 3, 64, dr));
  for (int i = 0; i < 6; ++i) EXPECT(dr[i] >= 0.0 && dr[i] <= 1.0);

  /* ---- Mcmc.rjmcmc_array with Interp jumps: evidence ratio 4.0 +- 0.1 (test/mcmc_test.ml:170-182) ----------------- */
  mg_rj_model A, B;
  memset(&A, 0, sizeof A); memset(&B, 0, sizeof B);
  A.like = like1; A.prior = prior; A.prop = prop; A.into.kind = MG_INTO_INTERP; A.into.tree = t1; A.p = 0.5;
  B.like = like2; B.prior = prior; B.prop = prop; B.into.kind = MG_INTO_INTERP; B.into.tree = t2; B.p = 0.5;
  mg_rjmcmc_cfg rcfg;
  memset(&rcfg, 0, sizeof rcfg);
  rcfg.nchains = 4096; rcfg.nbin = 50; rcfg.nskip = 10; rcfg.n = 250;
  uint8_t *model = malloc((size_t)rcfg.n * rcfg.nchains);
  int64_t counts[2] = {0, 0};
  CHECK(ctx, mg_rjmcmc_array(ctx, &A, &B, &rcfg, x0, x0, model, NULL, counts));
  EXPECT(counts[0] + counts[1] == rcfg.n * rcfg.nchains);
  const double ratio = (double)counts[0] / (double)counts[1];
  EXPECT(fabs(ratio - 4.0) < 0.1);
  int64_t jp_prop = 0, jp_acc = 0;
  CHECK(ctx, mg_rjmcmc_jump_counters(ctx, &jp_prop, &jp_acc));
  EXPECT(jp_prop > 0 && jp_acc > 0 && jp_acc <= jp_prop);

  /* ---- Evidence on a chain of a normalised 2-D Gaussian in the unit box: all three estimators ~ 1
   *      (evidence_test.ml:52-81: direct +- 0.5, Lebesgue +- 0.5; the harmonic mean is noisier) ---------------------- */
  double gl[4] = {0.5, 0.5, 0.1, 0.1};
  mg_logfn glike = {MG_FN_GAUSS_DIAG, D, 1.0, gl, 4};
  double gwrap[6] = {0, 0, 1, 1, 0.3, 0.3};
  mg_proposal gprop = {MG_PROP_WRAP, D, gwrap, 6};
  double *s3 = malloc(sizeof(double) * C * n * (D + 2));
  CHECK(ctx, mg_mcmc_array(ctx, &glike, &prior, &gprop, &cfg, x0, s3, NULL, NULL));
  const int64_t N = C * n;
  double *p3 = malloc(sizeof(double) * N * D), *ll = malloc(sizeof(double) * N), *lp = malloc(sizeof(double) * N);
  for (int64_t i = 0; i < N; ++i) { p3[i * D] = s3[i * (D + 2)]; p3[i * D + 1] = s3[i * (D + 2) + 1]; ll[i] = s3[i * (D + 2) + 2]; lp[i] = s3[i * (D + 2) + 3]; }
  double zh = 0, zl = 0, zd = 0;
  CHECK(ctx, mg_evidence_harmonic_mean(ctx, ll, N, &zh));
  CHECK(ctx, mg_evidence_lebesgue(ctx, p3, ll, lp, N, D, 64, 0.1, &zl));
  CHECK(ctx, mg_evidence_direct(ctx, p3, ll, lp, N, D, 64, &zd));
  EXPECT(zh > 0.1 && zh < 3.0 && fabs(zl - 1.0) < 0.5 && fabs(zd - 1.0) < 0.5);

  /* ---- Stats ------------------------------------------------------------------------------------------------------- */
  double mean[2], sd[2];
  CHECK(ctx, mg_stats_multi_mean(ctx, p1, N, D, mean));
  CHECK(ctx, mg_stats_multi_std(ctx, p1, N, D, mean, sd));
  EXPECT(fabs(mean[0] - 0.5) < 0.02 && fabs(sd[0] - sqrt(1.0 / 12.0)) < 0.01);
  double *g = malloc(sizeof(double) * 100000), gm = 0, gs = 0;
  CHECK(ctx, mg_stats_draw(ctx, MG_DRAW_GAUSSIAN, 3.0, 2.0, 100000, g));
  CHECK(ctx, mg_stats_mean(ctx, g, 100000, &gm));
  CHECK(ctx, mg_stats_std(ctx, g, 100000, 1, gm, &gs));
  EXPECT(fabs(gm - 3.0) < 0.05 && fabs(gs - 2.0) < 0.05);

  /* ---- Nested.nested_evidence: one normalised 2-D Gaussian, evidence 1 within 2x its error (nested_test.ml:23-39) - */
  double gpar[4] = {0.5, 0.5, 0.05, 0.05}, obox[5] = {0, 0, 1, 1, 0.0};
  mg_logfn nlike = {MG_FN_GAUSS_DIAG, D, 1.0, gpar, 4}, nprior = {MG_FN_BOX_OPEN, D, 1.0, obox, 5};
  mg_nested_cfg ncfg;
  memset(&ncfg, 0, sizeof ncfg);
  ncfg.dim = D; ncfg.nlive = 1000; ncfg.nmcmc = 100; ncfg.batch = 50; ncfg.epsrel = 0.01; ncfg.mode_hopping_frac = 0.1;
  ncfg.max_points = 400000;
  double *npts = malloc(sizeof(double) * ncfg.max_points * D), *nll = malloc(sizeof(double) * ncfg.max_points),
         *nlp = malloc(sizeof(double) * ncfg.max_points), *nlw = malloc(sizeof(double) * ncfg.max_points);
  double log_ev = 0, log_dev = 0; int64_t np = 0;
  CHECK(ctx, mg_nested_evidence(ctx, &nlike, &nprior, lo, hi, &ncfg, &log_ev, &log_dev, &np, npts, nll, nlp, nlw));
  const double err = exp(mg_nested_log_total_error(log_ev, log_dev, ncfg.nlive));
  EXPECT(np > ncfg.nlive && fabs(exp(log_ev) - 1.0) < 2.0 * err + 0.05);
  double wsum = 0; for (int64_t i = 0; i < np; ++i) wsum += exp(nlw[i]);
  EXPECT(fabs(wsum - 1.0) < 1e-8);                                            /* nested_test.ml:66-85 */
  for (int64_t i = 1; i < np; ++i) EXPECT(nll[i] >= nll[i - 1]);
  int64_t idx[1000];
  CHECK(ctx, mg_nested_posterior_indices(ctx, nlw, np, 1000, idx));
  double pm = 0; for (int i = 0; i < 1000; ++i) { EXPECT(idx[i] >= 0 && idx[i] < np); pm += npts[idx[i] * D]; }
  EXPECT(fabs(pm / 1000 - 0.5) < 0.02);

  /* ---- mcmcgpu_rjmcmc_array_k_native -> mg_rjmcmc_array_k: with two models it is the two-model call, draw for draw;
   *      three top hats (side 1, 1/2, 1/4: evidences 1, 1/4, 1/16), priors 0.2 / 0.3 / 0.5: time in model k ~ p_k Z_k ------ */
  {
    A.p = 0.5; B.p = 0.5;
    mg_rj_model two[2] = {A, B};
    const double *st2[2] = {x0, x0};
    int64_t c2[2] = {0, 0};
    CHECK(ctx, mg_ctx_set_seed(ctx, 77ull));
    CHECK(ctx, mg_rjmcmc_array(ctx, &A, &B, &rcfg, x0, x0, NULL, NULL, counts));
    CHECK(ctx, mg_ctx_set_seed(ctx, 77ull));
    CHECK(ctx, mg_rjmcmc_array_k(ctx, two, 2, &rcfg, st2, NULL, NULL, c2));
    EXPECT(c2[0] == counts[0] && c2[1] == counts[1]);
    double box3[5] = {0.375, 0.375, 0.625, 0.625, 0.0};
    mg_logfn like3 = {MG_FN_BOX_CLOSED, D, 1.0, box3, 5};
    double *s4 = malloc(sizeof(double) * C * n * (D + 2)), *p4 = malloc(sizeof(double) * C * n * D);
    CHECK(ctx, mg_mcmc_array(ctx, &like3, &prior, &prop, &cfg, x0, s4, NULL, NULL));
    for (int64_t i = 0; i < C * n; ++i) for (int d = 0; d < D; ++d) p4[i * D + d] = s4[i * (D + 2) + d];
    mg_kdtree *t3 = NULL;
    CHECK(ctx, mg_kdtree_build(ctx, p4, C * n, D, lo, hi, 2, &t3));
    mg_rj_model three[3] = {A, B, B};
    three[2].like = like3; three[2].into.tree = t3;
    three[0].p = 0.2; three[1].p = 0.3; three[2].p = 0.5;
    const double *st3[3] = {x0, x0, x0};
    int64_t c3[3] = {0, 0, 0};
    CHECK(ctx, mg_rjmcmc_array_k(ctx, three, 3, &rcfg, st3, NULL, NULL, c3));
    const double w0 = 0.2, w1 = 0.3 * 0.25, w2 = 0.5 * 0.0625, tot = (double)(c3[0] + c3[1] + c3[2]);
    EXPECT(c3[0] + c3[1] + c3[2] == rcfg.n * rcfg.nchains);
    EXPECT(fabs(c3[0] / tot - w0 / (w0 + w1 + w2)) < 0.02 && fabs(c3[1] / tot - w1 / (w0 + w1 + w2)) < 0.02 &&
           fabs(c3[2] / tot - w2 / (w0 + w1 + w2)) < 0.02);
    mg_kdtree_destroy(t3); free(s4); free(p4);
  }

  /* ---- mcmcgpu_enclosing_ellipse -> mg_ellipse_enclosing, mg_ellipse_range, mg_ellipse_tree_* (ellipse_test.ml:73-84:
   *      every point of the cloud lies inside enclosing_ellipse 2.0) ---------------------------------------------------- */
  {
    double cen[2], axes[2], ori[4], *rr = malloc(sizeof(double) * N);
    /* (the chains on the unit square accept every proposal, so no point repeats -- except slot 0 of every chain, the
     *  shared start point: a node of identical points makes the reference recurse forever and this library return
     *  MG_EFAIL, checked below; the cloud handed to the tree leaves those rows out) */
    double *p5 = malloc(sizeof(double) * N * D); int64_t N5 = 0;
    for (int64_t i = 0; i < N; ++i) if (i % n != 0) { p5[N5 * D] = p1[i * D]; p5[N5 * D + 1] = p1[i * D + 1]; ++N5; }
    CHECK(ctx, mg_ellipse_enclosing(ctx, p1, N, D, 2.0, cen, axes, ori));
    CHECK(ctx, mg_ellipse_range(ctx, cen, axes, ori, D, p1, N, rr));
    double rmax = 0; for (int64_t i = 0; i < N; ++i) { EXPECT(rr[i] < 1.0); rmax = rr[i] > rmax ? rr[i] : rmax; }
    EXPECT(fabs(rmax - 1.0 / sqrt(2.0)) < 1e-9 && fabs(cen[0] - 0.5) < 0.02 && axes[0] <= axes[1]);
    mg_ellipse_tree *et = NULL;
    EXPECT(mg_ellipse_tree_build(ctx, p1, N, D, 2.0, &et) == MG_EFAIL && et == NULL);   /* four copies of the start point */
    CHECK(ctx, mg_ellipse_tree_build(ctx, p5, N5, D, 2.0, &et));
    int64_t enp = 0, enn = 0; int32_t ed = 0, enl = 0;
    CHECK(ctx, mg_ellipse_tree_info(et, &enp, &ed, &enn, &enl));
    EXPECT(enp == N5 && ed == D && enn >= 1 && enl >= 1);
    int32_t *el = malloc(4 * enn), *er = malloc(4 * enn), *eb = malloc(4 * enn), *ee = malloc(4 * enn);
    CHECK(ctx, mg_ellipse_tree_export(et, el, er, eb, ee, NULL, NULL, NULL, NULL, NULL, NULL));
    EXPECT(eb[0] == 0 && ee[0] == N5);
    for (int64_t k = 0; k < enn; ++k) EXPECT(ee[k] - eb[k] >= D + 1 && el[k] < enn && er[k] < enn);
    mg_ellipse_tree_destroy(et); free(el); free(er); free(eb); free(ee); free(rr); free(p5);
  }

  /* ---- the pool: trim, reserve, same results ----------------------------------------------------------------------------- */
  {
    double zh2 = 0;
    CHECK(ctx, mg_ctx_trim_pool(ctx));
    CHECK(ctx, mg_ctx_reserve_pool(ctx, (int64_t)1 << 30));
    CHECK(ctx, mg_evidence_harmonic_mean(ctx, ll, N, &zh2));
    EXPECT(zh2 == zh);
  }

  /* ---- errors map as the stubs expect: MG_EINVAL <-> Invalid_argument, message available ---------------------------- */
  mg_mcmc_cfg bad = cfg; bad.nskip = 0;
  EXPECT(mg_mcmc_array(ctx, &like1, &prior, &prop, &bad, x0, s1, NULL, NULL) == MG_EINVAL && strlen(mg_last_error(ctx)) > 0);
  A.p = 0.7; B.p = 0.6;                                                      /* assert (pa + pb - 1 < sqrt eps), mcmc.ml:90 */
  EXPECT(mg_rjmcmc_array(ctx, &A, &B, &rcfg, x0, x0, NULL, NULL, counts) == MG_EFAIL);

  mg_kdtree_destroy(t1); mg_kdtree_destroy(t2);
  mg_ctx_destroy(ctx);
  free(s1); free(s2); free(s3); free(p1); free(p2); free(p3); free(model); free(ll); free(lp); free(g); free(npts); free(nll); free(nlp); free(nlw);
  printf("ok: mcmc_array, Interp, rjmcmc_array (ratio %.3f), Evidence (%.3f %.3f %.3f), Stats, nested_evidence (Z = %.4f +- %.4f)\n",
         ratio, zh, zl, zd, exp(log_ev), err);
  return 0;
}
