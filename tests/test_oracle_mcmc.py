"""Pin the oracle's mcmc.ml restatement against test/mcmc_test.ml (statistical
expectations, scaled down in sample count to keep the CPU suite short)."""
import numpy as np
import pytest

from mcmc_ocaml_b200 import plugins as P


def test_gaussian_post_uniform_proposal(og):  # mcmc_test.ml:40-59
    mu, sigma, n = 0.37, 1.62, 100000
    out, _, _ = og.mcmc_array(1, 0, n, P.gauss_diag([mu], [sigma]), P.zero(1), P.box_proposal([sigma]), [mu])
    x = out[:, 0, 0]
    tol = 10.0 * sigma / np.sqrt(n)
    assert abs(x.mean() - mu) < tol and abs(x.std() - sigma) < tol


def test_gaussian_post_left_biased_proposal(og):  # mcmc_test.ml:61-84: Hastings ratio
    mu, sigma, n = 0.61, 1.3, 100000
    out, _, _ = og.mcmc_array(2, 0, n, P.gauss_diag([mu], [sigma]), P.zero(1), P.left_biased_proposal(sigma), [mu])
    x = out[:, 0, 0]
    tol = 20.0 * sigma / np.sqrt(n)
    assert abs(x.mean() - mu) < tol and abs(x.std() - sigma) < tol


def test_prior_like_and_remove_repeat(og):  # mcmc_test.ml:86-112
    mu, sigma = 0.5, 1.5
    g = P.gauss_diag([mu], [sigma])
    out, _, _ = og.mcmc_array(3, 0, 20000, g.scaled(0.75), g.scaled(0.25), P.box_proposal([sigma]), [mu], nskip=10)
    x = out[:, 0, 0]
    assert abs(x.mean() - mu) < 0.2 * mu and abs(x.std() - sigma) < 0.2 * sigma
    rows = np.ascontiguousarray(out[:1000, :, 0])
    nr = og.remove_repeat_samples(rows, 1)
    assert np.all(nr[1:, 0] != nr[:-1, 0])


def test_rjmcmc_gaussians(og):  # mcmc_test.ml:114-148
    mu1, s1, mu2, s2 = 0.31, 0.62, 0.77, 0.45
    g1, g2 = P.gauss_diag([mu1], [s1]), P.gauss_diag([mu2], [s2])
    A = og.rj_model(g1.scaled(0.5), g1.scaled(0.5), P.indep_gauss_proposal([mu1], [s1]), 0.1, into_gauss=([mu1], [s1]))
    B = og.rj_model(g2.scaled(0.3), g2.scaled(0.7), P.indep_gauss_proposal([mu2], [s2]), 0.9, into_gauss=([mu2], [s2]))
    r = og.rjmcmc_array(4, 0, 2000, A, B, [mu1], [mu2], nskip=10, nchains=64, nthreads=8)
    n1, n2 = r["counts"]
    assert n1 / (n1 + n2) == pytest.approx(0.1, rel=0.1) and n2 / (n1 + n2) == pytest.approx(0.9, rel=0.1)
    assert n1 / n2 == pytest.approx(0.1 / 0.9, abs=0.1)


def test_rjmcmc_top_hats_interp(og):  # mcmc_test.ml:150-182: ratio 4.0 +- 0.1
    prior = P.box([0, 0], [1, 1], 0.0)
    like1, like2 = P.box([0, 0], [1, 1], 0.0), P.box([0.25, 0.25], [0.75, 0.75], 0.0)
    prop = P.wrap_proposal([0, 0], [1, 1], [0.5, 0.5])
    s1, _, _ = og.mcmc_array(5, 0, 10000, like1, prior, prop, [0.5, 0.5], nskip=20)
    s2, _, _ = og.mcmc_array(5, 1, 10000, like2, prior, prop, [0.5, 0.5], nskip=20)
    t1 = og.Tree(np.ascontiguousarray(s1[:, :2, 0]), [0, 0], [1, 1])
    t2 = og.Tree(np.ascontiguousarray(s2[:, :2, 0]), [0, 0], [1, 1])
    A, B = og.rj_model(like1, prior, prop, 0.5, tree=t1), og.rj_model(like2, prior, prop, 0.5, tree=t2)
    r = og.rjmcmc_array(6, 0, 4000, A, B, [0.5, 0.5], [0.5, 0.5], nskip=10, nchains=64, nthreads=8, record_model=False)
    assert r["counts"][0] / r["counts"][1] == pytest.approx(4.0, abs=0.15)


def test_combine_jump_proposal(og):  # mcmc_test.ml:184-208
    prop = P.combine_jump_proposals([(1.0, P.one_sided_proposal(-1.0)), (2.0, P.one_sided_proposal(+1.0))])
    out, _, _ = og.mcmc_array(7, 0, 4000, P.gauss_diag([0.0], [1.0]), P.zero(1), prop, [0.0], nskip=5, nchains=64, nthreads=8)
    x = out[:, 0, :].ravel()
    assert abs(x.mean()) < 0.05 and x.std(ddof=1) == pytest.approx(1.0, rel=2e-2)


# ---- k-model reversible jump: an extension, specified in oracle.cpp: rj_chain_k ------------------

def _three_gaussians(og):
    """Three normalised 1-D Gaussians split differently between likelihood and prior: every model has evidence 1, so
    the chain visits model k with its prior probability (the k-model form of mcmc_test.ml:114-148)."""
    spec = [(0.31, 0.62, 0.5, 0.2), (0.77, 0.45, 0.3, 0.3), (-0.4, 0.8, 0.6, 0.5)]
    ms = []
    for mu, s, w, p in spec:
        g = P.gauss_diag([mu], [s])
        ms.append(og.rj_model(g.scaled(w), g.scaled(1.0 - w), P.indep_gauss_proposal([mu], [s]), p, into_gauss=([mu], [s])))
    return ms, [[m] for m, _, _, _ in spec], [p for _, _, _, p in spec]


def test_rjmcmc_k_with_two_models_is_the_two_model_sampler(og):
    mu1, s1, mu2, s2 = 0.31, 0.62, 0.77, 0.45
    g1, g2 = P.gauss_diag([mu1], [s1]), P.gauss_diag([mu2], [s2])
    A = og.rj_model(g1.scaled(0.5), g1.scaled(0.5), P.indep_gauss_proposal([mu1], [s1]), 0.1, into_gauss=([mu1], [s1]))
    B = og.rj_model(g2.scaled(0.3), g2.scaled(0.7), P.indep_gauss_proposal([mu2], [s2]), 0.9, into_gauss=([mu2], [s2]))
    two = og.rjmcmc_array(11, 3, 200, A, B, [mu1], [mu2], nskip=3, nbin=7, nchains=32, nthreads=4, record_samples=True)
    k = og.rjmcmc_array_k(11, 3, 200, [A, B], [[mu1], [mu2]], nskip=3, nbin=7, nchains=32, nthreads=4, record_samples=True)
    assert np.array_equal(two["model"], k["model"]) and np.array_equal(two["samples"], k["samples"])
    assert two["counts"] == k["counts"] and two["accept"] == k["accept"] and two["cross"] == k["cross"]


def test_rjmcmc_k_three_models_recover_their_priors(og):
    ms, starts, priors = _three_gaussians(og)
    r = og.rjmcmc_array_k(4, 0, 2000, ms, starts, nskip=10, nchains=64, nthreads=8)
    frac = np.array(r["counts"]) / sum(r["counts"])
    assert sum(r["counts"]) == 2000 * 64
    np.testing.assert_allclose(frac, priors, rtol=0.1)
    assert r["model"].max() == 2 and r["cross"][0] > 0


def test_rjmcmc_k_prior_sum_assertion(og):
    ms, starts, _ = _three_gaussians(og)
    ms[2] = (type(ms[2][0])(ms[2][0].like, ms[2][0].prior, ms[2][0].prop, ms[2][0].into, 0.6), ms[2][1])   # 0.2 + 0.3 + 0.6
    with pytest.raises(Exception):
        og.rjmcmc_array_k(4, 0, 10, ms, starts)


def test_rjmcmc_k_three_top_hats_known_evidence_ratios(og):
    """The k-model form of mcmc_test.ml:150-182: top hats of side 1, 1/2, 1/4 on the unit square (evidences 1, 1/4, 1/16),
    interpolated jumps into each, model priors (0.2, 0.3, 0.5): time in model k is proportional to p_k Z_k."""
    prior = P.box([0, 0], [1, 1], 0.0)
    boxes = [([0, 0], [1, 1]), ([0.25, 0.25], [0.75, 0.75]), ([0.375, 0.375], [0.625, 0.625])]
    prop = P.wrap_proposal([0, 0], [1, 1], [0.5, 0.5])
    pri = [0.2, 0.3, 0.5]
    ms = []
    for k, ((lo, hi), p) in enumerate(zip(boxes, pri)):
        like = P.box(lo, hi, 0.0)
        s, _, _ = og.mcmc_array(5, k, 6000, like, prior, prop, [0.5, 0.5], nskip=20)
        ms.append(og.rj_model(like, prior, prop, p, tree=og.Tree(np.ascontiguousarray(s[:, :2, 0]), [0, 0], [1, 1])))
    r = og.rjmcmc_array_k(6, 0, 3000, ms, [[0.5, 0.5]] * 3, nskip=10, nbin=100, nchains=64, nthreads=8, record_model=False)
    w = np.array(pri) * np.array([1.0, 0.25, 0.0625])
    np.testing.assert_allclose(np.array(r["counts"]) / sum(r["counts"]), w / w.sum(), rtol=0.08)
