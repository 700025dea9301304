"""Reversible-jump ensembles (mg_rjmcmc_array) against the oracle and against
the known answers of the reference's tests (test/mcmc_test.ml:114-182)."""
import numpy as np
import pytest
from scipy import stats

from mcmc_ocaml_b200 import Failure, interpolate_pdf, mcmc, plugins as P
from tests._parity import divergence_report

pytestmark = pytest.mark.gpu


def tophat_setup(ctx, og, nsamp=10000):
    """test/mcmc_test.ml:150-182: unit square vs the central 0.5 x 0.5 square"""
    prior = P.box([0, 0], [1, 1], 0.0)
    like1 = P.box([0, 0], [1, 1], 0.0)
    like2 = P.box([0.25, 0.25], [0.75, 0.75], 0.0)
    prop = P.wrap_proposal([0, 0], [1, 1], [0.5, 0.5])
    ctx.set_seed(1001)
    s1 = mcmc.mcmc_array(nsamp, like1, prior, prop, [0.5, 0.5], nskip=100, ctx=ctx)
    s2 = mcmc.mcmc_array(nsamp, like2, prior, prop, [0.5, 0.5], nskip=100, ctx=ctx)
    p1, p2 = s1.values(), s2.values()
    return prior, like1, like2, prop, p1, p2


def test_tophat_interp_matches_oracle_and_known_ratio(ctx, og):
    prior, like1, like2, prop, p1, p2 = tophat_setup(ctx, og)
    i1 = interpolate_pdf.InterpPdf(p1, [0, 0], [1, 1], ctx=ctx)
    i2 = interpolate_pdf.InterpPdf(p2, [0, 0], [1, 1], ctx=ctx)
    A = mcmc.RjModel(like1, prior, prop, 0.5, interp=i1)
    B = mcmc.RjModel(like2, prior, prop, 0.5, interp=i2)
    # (1) chain-for-chain against the oracle on the same Philox stream
    ctx.set_seed(77)
    g = mcmc.rjmcmc_array(300, A, B, [0.5, 0.5], [0.5, 0.5], nskip=3, nbin=10, nchains=256, record_samples=True, ctx=ctx)
    t1, t2 = og.Tree(p1, [0, 0], [1, 1]), og.Tree(p2, [0, 0], [1, 1])
    oa = og.rj_model(like1, prior, prop, 0.5, tree=t1)
    ob = og.rj_model(like2, prior, prop, 0.5, tree=t2)
    o = og.rjmcmc_array(77, 0, 300, oa, ob, [0.5, 0.5], [0.5, 0.5], nskip=3, nbin=10, nchains=256, nthreads=8,
                        record_samples=True, margins=True)
    # log(jump_prob): CUDA libm vs glibc may flip a decision only at a near-tie of the accept test
    assert divergence_report(g.block[:, :2, :], o["samples"][:, :2, :], o["margins"], "RJ tophat") <= 0.01
    same = np.all(g.model == o["model"], axis=0)
    assert np.array_equal(g.block[:, :2, same], o["samples"][:, :2, same])
    assert abs(g.counts[0] - o["counts"][0]) <= 0.01 * 300 * 256
    # (2) the reference's known answer: evidence ratio 4.0 +- 0.1 (mcmc_test.ml:181-182),
    # here with 4096 chains x 250 samples at nskip = 10 (1.02e6 samples, as the test's 1e6)
    ctx.set_seed(78)
    r = mcmc.rjmcmc_array(250, A, B, [0.5, 0.5], [0.5, 0.5], nskip=10, nbin=50, nchains=4096, record_model=False, ctx=ctx)
    assert mcmc.rjmcmc_evidence_ratio(r) == pytest.approx(4.0, abs=0.1)
    # high-level jumps (draw_high_level / jump_prob_high_level) give the same answer
    A64 = mcmc.RjModel(like1, prior, prop, 0.5, interp=i1, nstop=64)
    B64 = mcmc.RjModel(like2, prior, prop, 0.5, interp=i2, nstop=64)
    r = mcmc.rjmcmc_array(250, A64, B64, [0.5, 0.5], [0.5, 0.5], nskip=10, nbin=50, nchains=4096, record_model=False, ctx=ctx)
    assert mcmc.rjmcmc_evidence_ratio(r) == pytest.approx(4.0, abs=0.15)


def test_rjmcmc_gaussians_priors_recovered(ctx, og):
    """test/mcmc_test.ml:114-148: two normalised 1-D Gaussians, priors (0.1, 0.9)"""
    mu1, s1, mu2, s2 = 0.31, 0.62, 0.77, 0.45
    g1, g2 = P.gauss_diag([mu1], [s1]), P.gauss_diag([mu2], [s2])
    A = mcmc.RjModel(g1.scaled(0.5), g1.scaled(0.5), P.indep_gauss_proposal([mu1], [s1]), 0.1, into_gauss=([mu1], [s1]))
    B = mcmc.RjModel(g2.scaled(0.3), g2.scaled(0.7), P.indep_gauss_proposal([mu2], [s2]), 0.9, into_gauss=([mu2], [s2]))
    ctx.set_seed(5)
    r = mcmc.rjmcmc_array(250, A, B, [mu1], [mu2], nskip=10, nchains=4096, ctx=ctx)
    n1, n2 = r.counts
    pp1, pp2 = n1 / (n1 + n2), n2 / (n1 + n2)
    assert pp1 == pytest.approx(0.1, rel=0.1) and pp2 == pytest.approx(0.9, rel=0.1)
    assert mcmc.rjmcmc_evidence_ratio(r) == pytest.approx(0.1 / 0.9, abs=0.1)
    # binomial test of the model fraction over (nearly independent) chains at the last sample
    last = r.model[-1]
    assert stats.binomtest(int((last == 0).sum()), last.size, 0.1).pvalue > 1e-4
    # same stream in the oracle: model sequences agree chain for chain
    oa = og.rj_model(g1.scaled(0.5), g1.scaled(0.5), P.indep_gauss_proposal([mu1], [s1]), 0.1, into_gauss=([mu1], [s1]))
    ob = og.rj_model(g2.scaled(0.3), g2.scaled(0.7), P.indep_gauss_proposal([mu2], [s2]), 0.9, into_gauss=([mu2], [s2]))
    ctx.set_seed(6)
    g = mcmc.rjmcmc_array(100, A, B, [mu1], [mu2], nskip=2, nchains=128, ctx=ctx)
    o = og.rjmcmc_array(6, 0, 100, oa, ob, [mu1], [mu2], nskip=2, nchains=128, nthreads=8, margins=True)
    assert divergence_report(g.model, o["model"], o["margins"], "RJ two Gaussians") <= 0.02


def test_gaussian_vs_cauchy_config1(ctx, og):
    """BASELINE.json config 1 at test size: Gaussian-vs-Cauchy on the fixed
    100-point dataset, interpolated jumps (bin/gaussian_cauchy_efficiency.ml)."""
    from tests.golden.gc_data import DATA
    prior = P.box([-1.0, 0.5], [1.0, 1.5], value=-0.693147)
    prop = P.wrap_proposal([-1.0, 0.5], [1.0, 1.5], [0.1, 0.1])
    lg, lc = P.gauss_data(DATA), P.cauchy_data(DATA)
    ctx.set_seed(20111104)
    gs = mcmc.mcmc_array(200, lg, prior, prop, [0.0, 1.0], nchains=64, nbin=500, nskip=20, ctx=ctx).values()
    cs = mcmc.mcmc_array(200, lc, prior, prop, [0.0, 1.0], nchains=64, nbin=500, nskip=20, ctx=ctx).values()
    lo, hi = [-1.0, 0.5], [1.0, 1.5]
    gi = interpolate_pdf.InterpPdf(gs, lo, hi, ctx=ctx)
    ci = interpolate_pdf.InterpPdf(cs, lo, hi, ctx=ctx)
    A = mcmc.RjModel(lg, prior, prop, 0.5, interp=gi)
    B = mcmc.RjModel(lc, prior, prop, 0.5, interp=ci)
    ctx.set_seed(9)
    g = mcmc.rjmcmc_array(400, A, B, [0.0, 1.0], [0.0, 1.0], nskip=5, nbin=100, nchains=2048, record_model=False, ctx=ctx)
    oa = og.rj_model(lg, prior, prop, 0.5, tree=og.Tree(gs, lo, hi))
    ob = og.rj_model(lc, prior, prop, 0.5, tree=og.Tree(cs, lo, hi))
    o = og.rjmcmc_array(10, 0, 400, oa, ob, [0.0, 1.0], [0.0, 1.0], nskip=5, nbin=100, nchains=512, nthreads=8,
                        record_model=False)
    fg = g.counts[0] / sum(g.counts)
    fo = o["counts"][0] / sum(o["counts"])
    # the data were drawn from a Gaussian: the Gaussian model must dominate, and
    # GPU and oracle (independent streams) must agree within Monte Carlo error
    assert fg > 0.8
    assert abs(fg - fo) < 0.03


def test_prior_assertion(ctx):
    g = P.gauss_diag([0.0], [1.0])
    A = mcmc.RjModel(g, P.zero(1), P.box_proposal([1.0]), 0.7, into_gauss=([0.0], [1.0]))
    B = mcmc.RjModel(g, P.zero(1), P.box_proposal([1.0]), 0.6, into_gauss=([0.0], [1.0]))
    with pytest.raises(Failure):          # assert (pa +. pb -. 1.0 < sqrt epsilon_float), mcmc.ml:90
        mcmc.rjmcmc_array(10, A, B, [0.0], [0.0], ctx=ctx)


@pytest.mark.parametrize("dA,dB,nstop", [(8, 16, 0), (8, 16, 64), (3, 5, 0)])
def test_high_dimensional_interp_jumps_match_oracle(ctx, og, dA, dB, nstop):
    """BASELINE.json config 5's model pair at test size: isotropic Gaussian posteriors of dA and dB dimensions
    (flat priors of density 1 and 1/2 on the unit box, so Z_A / Z_B = 2), trees of 1e5 posterior draws each,
    interpolated jumps at leaf level (nstop = 0) and through the *_high_level forms.  Chain for chain against the
    oracle on the same Philox stream; a chain may leave the oracle only at a near-tie of the accept test."""
    import math
    s, ntree = 0.05, 100_000
    rng = np.random.default_rng(dA * 100 + dB)
    gm, om = [], []
    for d, logc in ((dA, 0.0), (dB, -math.log(2.0))):
        pts = rng.normal(0.5, s, (ntree, d)).clip(0.0, 1.0)
        lo, hi = np.zeros(d), np.ones(d)
        like = P.gauss_diag(np.full(d, 0.5), np.full(d, s))
        prior = P.box(lo, hi, logc)
        prop = P.wrap_proposal(lo, hi, np.full(d, 2.0 * s / math.sqrt(d)))
        gm.append(mcmc.RjModel(like, prior, prop, 0.5, interp=interpolate_pdf.InterpPdf(pts, lo, hi, ctx=ctx), nstop=nstop))
        om.append(og.rj_model(like, prior, prop, 0.5, tree=og.Tree(pts, lo, hi), nstop=nstop))
    a0, b0 = np.full(dA, 0.5), np.full(dB, 0.5)
    C, n = 512, 120
    ctx.set_seed(4242)
    g = mcmc.rjmcmc_array(n, gm[0], gm[1], a0, b0, nskip=2, nbin=5, nchains=C, record_samples=True, ctx=ctx)
    o = og.rjmcmc_array(4242, 0, n, om[0], om[1], a0, b0, nskip=2, nbin=5, nchains=C, nthreads=8, record_samples=True,
                        margins=True)
    Dm = max(dA, dB)
    frac = divergence_report(np.concatenate([g.block[:, :Dm, :], g.model[:, None, :].astype(float)], axis=1),
                             np.concatenate([o["samples"][:, :Dm, :], o["model"][:, None, :].astype(float)], axis=1),
                             o["margins"], f"RJ ({dA},{dB})-D nstop={nstop}")
    assert frac <= 0.01
    same = np.all(g.model == o["model"], axis=0)
    np.testing.assert_allclose(g.block[:, Dm, same], o["samples"][:, Dm, same], rtol=1e-12, atol=1e-12)   # ll
    if frac == 0.0:
        assert g.counts == o["counts"] and g.cross == o["cross"]
    assert g.cross[0] > 0.4 * C * (n - 1) * 2          # half of the steps propose a jump into the other model


# ---- k-model reversible jump (mg_rjmcmc_array_k): an extension, SURVEY 8f rank 3 -----------------------------------

def test_k_model_call_with_two_models_equals_the_two_model_call(ctx):
    prior, like1, like2, prop, p1, p2 = tophat_setup(ctx, None, nsamp=4000)
    i1 = interpolate_pdf.InterpPdf(p1, [0, 0], [1, 1], ctx=ctx)
    i2 = interpolate_pdf.InterpPdf(p2, [0, 0], [1, 1], ctx=ctx)
    A = mcmc.RjModel(like1, prior, prop, 0.4, interp=i1)
    B = mcmc.RjModel(like2, prior, prop, 0.6, interp=i2)
    ctx.set_seed(501)
    two = mcmc.rjmcmc_array(150, A, B, [0.5, 0.5], [0.5, 0.5], nskip=3, nbin=10, nchains=333, record_samples=True, ctx=ctx)
    ctx.set_seed(501)
    k = mcmc.rjmcmc_array_k(150, [A, B], [[0.5, 0.5], [0.5, 0.5]], nskip=3, nbin=10, nchains=333, record_samples=True, ctx=ctx)
    assert np.array_equal(two.model, k.model) and np.array_equal(two.block, k.block)
    assert two.counts == k.counts and two.cross == k.cross


def test_three_models_match_oracle_and_known_evidence_ratios(ctx, og):
    """Three top hats on the unit square (sides 1, 1/2, 1/4: evidences 1, 1/4, 1/16 under the unit-square prior) with
    interpolated jumps into each, model priors (0.2, 0.3, 0.5): chain for chain against the oracle on the same Philox
    stream, then the known answer -- time in model k proportional to p_k Z_k."""
    prior = P.box([0, 0], [1, 1], 0.0)
    boxes = [([0, 0], [1, 1]), ([0.25, 0.25], [0.75, 0.75]), ([0.375, 0.375], [0.625, 0.625])]
    prop = P.wrap_proposal([0, 0], [1, 1], [0.5, 0.5])
    pri = [0.2, 0.3, 0.5]
    ctx.set_seed(1001)
    gm, om = [], []
    for (lo, hi), p in zip(boxes, pri):
        like = P.box(lo, hi, 0.0)
        pts = mcmc.mcmc_array(6000, like, prior, prop, [0.5, 0.5], nskip=50, ctx=ctx).values()
        gm.append(mcmc.RjModel(like, prior, prop, p, interp=interpolate_pdf.InterpPdf(pts, [0, 0], [1, 1], ctx=ctx)))
        om.append(og.rj_model(like, prior, prop, p, tree=og.Tree(pts, [0, 0], [1, 1])))
    starts = [[0.5, 0.5]] * 3
    ctx.set_seed(88)
    g = mcmc.rjmcmc_array_k(200, gm, starts, nskip=3, nbin=10, nchains=256, record_samples=True, ctx=ctx)
    o = og.rjmcmc_array_k(88, 0, 200, om, starts, nskip=3, nbin=10, nchains=256, nthreads=8, record_samples=True, margins=True)
    frac = divergence_report(np.concatenate([g.block[:, :2, :], g.model[:, None, :].astype(float)], axis=1),
                             np.concatenate([o["samples"][:, :2, :], o["model"][:, None, :].astype(float)], axis=1),
                             o["margins"], "RJ three top hats")
    assert frac <= 0.01
    if frac == 0.0:
        assert g.counts == o["counts"] and g.cross == o["cross"]
    assert g.model.max() == 2
    ctx.set_seed(89)
    r = mcmc.rjmcmc_array_k(250, gm, starts, nskip=10, nbin=100, nchains=4096, record_model=False, ctx=ctx)
    w = np.array(pri) * np.array([1.0, 0.25, 0.0625])
    got = np.array(r.counts) / sum(r.counts)
    np.testing.assert_allclose(got, w / w.sum(), rtol=0.05)
    assert r.counts[0] / r.counts[1] == pytest.approx((0.2 * 1.0) / (0.3 * 0.25), rel=0.05)


def test_k_model_arguments(ctx):
    g = P.gauss_diag([0.0], [1.0])
    mk = lambda p: mcmc.RjModel(g, P.zero(1), P.box_proposal([1.0]), p, into_gauss=([0.0], [1.0]))
    with pytest.raises(Failure):                       # sum of the priors, as mcmc.ml:90
        mcmc.rjmcmc_array_k(10, [mk(0.5), mk(0.4), mk(0.3)], [[0.0]] * 3, ctx=ctx)
    from mcmc_ocaml_b200 import InvalidArgument
    with pytest.raises(InvalidArgument):               # 2..MG_RJ_MAX_MODELS models
        mcmc.rjmcmc_array_k(10, [mk(0.1)] * 9, [[0.0]] * 9, ctx=ctx)
    r = mcmc.rjmcmc_array_k(50, [mk(0.125)] * 8, [[0.0]] * 8, nchains=512, ctx=ctx)
    assert sum(r.counts) == 50 * 512 and min(r.counts) > 0.08 * 50 * 512   # eight equal models: about 1/8 each


def test_four_models_of_different_dimensions_match_oracle(ctx, og):
    """k-model sampler across dimensions: isotropic Gaussian posteriors of 2, 3, 5 and 8 dimensions (evidences 1, 1/2,
    1/4, 1/8 through the prior's density), interpolated jumps at leaf level and through the *_high_level forms, unequal
    model priors.  Chain for chain against the oracle's rj_chain_k on the same Philox stream."""
    import math
    s, ntree = 0.05, 60_000
    dims, pri, nstops = (2, 3, 5, 8), (0.1, 0.2, 0.3, 0.4), (0, 0, 32, 0)
    rng = np.random.default_rng(2468)
    gm, om, starts = [], [], []
    for k, d in enumerate(dims):
        pts = rng.normal(0.5, s, (ntree, d)).clip(0.0, 1.0)
        lo, hi = np.zeros(d), np.ones(d)
        like = P.gauss_diag(np.full(d, 0.5), np.full(d, s))
        prior = P.box(lo, hi, -k * math.log(2.0))
        prop = P.wrap_proposal(lo, hi, np.full(d, 2.0 * s / math.sqrt(d)))
        gm.append(mcmc.RjModel(like, prior, prop, pri[k], interp=interpolate_pdf.InterpPdf(pts, lo, hi, ctx=ctx), nstop=nstops[k]))
        om.append(og.rj_model(like, prior, prop, pri[k], tree=og.Tree(pts, lo, hi), nstop=nstops[k]))
        starts.append(np.full(d, 0.5))
    C, n = 384, 100
    ctx.set_seed(1357)
    g = mcmc.rjmcmc_array_k(n, gm, starts, nskip=2, nbin=5, nchains=C, record_samples=True, ctx=ctx)
    o = og.rjmcmc_array_k(1357, 0, n, om, starts, nskip=2, nbin=5, nchains=C, nthreads=8, record_samples=True, margins=True)
    Dm = max(dims)
    frac = divergence_report(np.concatenate([g.block[:, :Dm, :], g.model[:, None, :].astype(float)], axis=1),
                             np.concatenate([o["samples"][:, :Dm, :], o["model"][:, None, :].astype(float)], axis=1),
                             o["margins"], "RJ four models (2,3,5,8)-D")
    assert frac <= 0.01
    if frac == 0.0:
        assert g.counts == o["counts"] and g.cross == o["cross"]
    assert set(np.unique(g.model)) == {0, 1, 2, 3} and g.cross[1] > 0
