"""Evidence estimators and Stats reductions on the GPU against the oracle
(deterministic evaluation: relative tolerance 1e-12)."""
import math

import numpy as np
import pytest

from mcmc_ocaml_b200 import InvalidArgument, evidence, mcmc, plugins as P, stats

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def mh_samples(ctx, seed, D, n, nchains=1, sigma=0.05):
    mu = np.full(D, 0.5)
    like = P.gauss_diag(mu, np.full(D, sigma))
    prior = P.box(np.zeros(D), np.ones(D), 0.0)
    prop = P.box_proposal(np.full(D, sigma / max(1.0, D / 2)))
    ctx.set_seed(seed)
    s = mcmc.mcmc_array(n, like, prior, prop, mu, nchains=nchains, ctx=ctx)
    return s.values(), s.log_likelihood(), s.log_prior()


def test_harmonic_mean(ctx, og):
    rng = np.random.default_rng(0)
    for n in (1, 2, 1000, 300001):
        ll = rng.normal(-3.0, 2.0, n)
        seq, ld = og.evidence_harmonic_mean(ll)
        got = evidence.evidence_harmonic_mean(ll=ll, ctx=ctx)
        assert got == pytest.approx(ld, rel=RTOL)          # vs the extended-precision sum
        assert got == pytest.approx(seq, rel=1e-11)         # vs the reference's left-to-right sum


@pytest.mark.parametrize("D,n,nb", [(2, 10000, 64), (2, 3000, 16), (5, 20000, 64), (20, 30000, 64), (1, 5000, 32), (40, 6000, 64)])
def test_lebesgue_matches_oracle(ctx, og, D, n, nb):
    pts, ll, lp = mh_samples(ctx, 100 + D, D, n)            # MH output: ~50 % exact repeats
    for eps in (0.1, 0.2, 1e9):
        want = og.evidence_lebesgue(pts, ll, lp, n=nb, eps=eps)
        got = evidence.evidence_lebesgue(pts, ll, lp, n=nb, eps=eps, ctx=ctx)
        assert got == pytest.approx(want["value_ld"], rel=RTOL)
        assert got == pytest.approx(want["value"], rel=1e-11)


@pytest.mark.parametrize("D,n,nb", [(2, 10000, 64), (3, 4000, 8), (10, 20000, 64)])
def test_direct_matches_oracle(ctx, og, D, n, nb):
    pts, ll, lp = mh_samples(ctx, 200 + D, D, n)
    want = og.evidence_direct(pts, ll, lp, n=nb)
    got = evidence.evidence_direct(pts, ll, lp, n=nb, ctx=ctx)
    assert got == pytest.approx(want["value_ld"], rel=RTOL)
    assert got == pytest.approx(want["value"], rel=1e-11)


def test_direct_lexicographic_fallback(ctx, og):
    """ties in coordinate 0 that are NOT whole repeated rows force the full lexicographic sort"""
    rng = np.random.default_rng(12)
    pts = np.round(rng.random((6000, 3)), 1)            # a coarse grid: many partial ties, many exact duplicates
    pts[:, 2] = rng.random(6000)
    pts[100:200] = pts[0:100]                           # and whole repeated rows
    ll = rng.normal(-2.0, 1.0, 6000); lp = rng.normal(-1.0, 0.3, 6000)
    ll[100:200] = ll[0:100]; lp[100:200] = lp[0:100]
    want = og.evidence_direct(pts, ll, lp, n=16)
    got = evidence.evidence_direct(pts, ll, lp, n=16, ctx=ctx)
    assert got == pytest.approx(want["value_ld"], rel=RTOL)


def test_lebesgue_pooled_chains_and_known_answer(ctx, og):
    """evidence_test.ml:74-81 (Lebesgue ~ 1 for a normalised Gaussian in the
    unit box), on 64 pooled chains"""
    pts, ll, lp = mh_samples(ctx, 7, 2, 400, nchains=64, sigma=0.06)
    got = evidence.evidence_lebesgue(pts, ll, lp, n=64, eps=0.2, ctx=ctx)
    assert got == pytest.approx(1.0, abs=0.5)
    assert got == pytest.approx(og.evidence_lebesgue(pts, ll, lp, n=64, eps=0.2)["value_ld"], rel=RTOL)
    d = evidence.evidence_direct(pts, ll, lp, n=64, ctx=ctx)
    assert d == pytest.approx(og.evidence_direct(pts, ll, lp, n=64)["value_ld"], rel=RTOL)
    h = evidence.evidence_harmonic_mean(ll=ll, ctx=ctx)
    assert h == pytest.approx(og.evidence_harmonic_mean(ll)[1], rel=RTOL)


def test_edge_cases(ctx, og):
    # all samples identical: one leaf of N >= n duplicates contributes nothing (evidence.ml:83-89)
    pts = np.repeat([[0.3, 0.4]], 100, axis=0); ll = np.full(100, -1.0); lp = np.zeros(100)
    assert evidence.evidence_direct(pts, ll, lp, n=64, ctx=ctx) == og.evidence_direct(pts, ll, lp, n=64)["value"]
    # fewer samples than n: the root is the only cell
    rng = np.random.default_rng(4)
    pts = rng.random((10, 3)); ll = rng.normal(size=10); lp = rng.normal(size=10)
    assert evidence.evidence_lebesgue(pts, ll, lp, n=64, eps=1e9, ctx=ctx) == pytest.approx(
        og.evidence_lebesgue(pts, ll, lp, n=64, eps=1e9)["value"], rel=RTOL)
    assert evidence.evidence_direct(pts, ll, lp, n=64, ctx=ctx) == pytest.approx(
        og.evidence_direct(pts, ll, lp, n=64)["value"], rel=RTOL)
    with pytest.raises(InvalidArgument):
        evidence.evidence_direct(np.zeros((0, 2)), np.zeros(0), np.zeros(0), ctx=ctx)   # bounds_of_objects []


def test_stats_goldens_and_oracle(ctx, og):
    # test/stats_test.ml goldens through the GPU path
    assert stats.mean([0.0, 1.0, 2.0, 3.0], ctx=ctx) == pytest.approx(1.5, rel=1e-15)
    assert stats.std([1.0, 2.0, 3.0, 4.0, 5.0], ctx=ctx) == pytest.approx(math.sqrt(10.0) / 2.0, rel=1e-15)
    np.testing.assert_allclose(stats.multi_mean([[0.0, 1.0], [2.0, 3.0], [4.0, -5.0]], ctx=ctx), [2.0, -1.0 / 3.0], rtol=1e-15)
    xs = [[0.662891, 0.218155, 0.464706, 0.148477, 0.39616], [0.43397, 0.161041, 0.625332, 0.508765, 0.261084],
          [0.147267, 0.403388, 0.643601, 0.892214, 0.269893]]
    np.testing.assert_allclose(stats.multi_std(xs, ctx=ctx), [0.258351, 0.126692, 0.098436, 0.371928, 0.0755716], rtol=1e-3, atol=1e-3)
    rng = np.random.default_rng(2)
    for n, d in [(100000, 1), (50001, 7), (20000, 20), (3000, 64)]:
        x = rng.normal(3.0, 2.0, (n, d))
        np.testing.assert_allclose(stats.multi_mean(x, ctx=ctx), og.multi_mean(x), rtol=RTOL)
        np.testing.assert_allclose(stats.multi_std(x, ctx=ctx), og.multi_std(x), rtol=RTOL)
        mu = og.multi_mean(x)
        np.testing.assert_allclose(stats.multi_std(x, mean=mu, ctx=ctx), og.multi_std(x, mean=mu), rtol=RTOL)
    x = rng.normal(size=20000)
    r, L = stats.autocorrelation(x, 50, ctx=ctx)
    ro, Lo = og.autocorrelation(x, 50)
    np.testing.assert_allclose(r, ro, rtol=1e-10, atol=1e-13)
    assert L == pytest.approx(Lo, rel=1e-10)


def test_harmonic_bootstrap_and_cli_tools(ctx, og, tmp_path):
    """bin/harmonic_evidence.ml (bootstrap) and bin/evidence_tool.ml equivalents, through the
    Read_write text format"""
    import subprocess
    import sys as _sys
    import os as _os
    from mcmc_ocaml_b200 import read_write
    rng = np.random.default_rng(5)
    ll = rng.normal(-1.0, 0.7, 4001)
    ctx.set_seed(9)
    evs = evidence.harmonic_bootstrap(ll, 300, ctx=ctx)
    want = np.sort(og.harmonic_bootstrap(9, 0, ll, 300))
    np.testing.assert_allclose(evs, want, rtol=1e-11)       # same resampled indices, compensated vs sequential sum
    pts, l2, lp = mh_samples(ctx, 41, 2, 3000)
    rows = np.concatenate([pts, l2[:, None], lp[:, None]], axis=1)
    f = str(tmp_path / "chain.dat")
    read_write.write(f, rows, lossless=True)
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    out = subprocess.run([_sys.executable, _os.path.join(root, "tools", "evidence_tool.py"), "-nbox", "32", "-lebeps", "0.2", "-i", f],
                         capture_output=True, text=True, check=True).stdout.split()
    h, l, d = (float(x) for x in out)
    assert h == pytest.approx(og.evidence_harmonic_mean(l2)[0], rel=1e-5)          # printed with %g
    assert l == pytest.approx(og.evidence_lebesgue(pts, l2, lp, n=32, eps=0.2)["value"], rel=1e-5)
    assert d == pytest.approx(og.evidence_direct(pts, l2, lp, n=32)["value"], rel=1e-5)
    out = subprocess.run([_sys.executable, _os.path.join(root, "tools", "harmonic_evidence.py"), "-nbstrap", "200", "-seed", "3", "-i", f],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    best, lo, hi = (float(x) for x in out[1].split())
    assert lo <= best * 1.5 and hi >= best * 0.5 and lo < hi
