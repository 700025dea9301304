"""Nested sampling with batched constrained replacement against the oracle
(same Philox stream: point-for-point) and the reference's known answers
(test/nested_test.ml)."""
import math

import numpy as np
import pytest

from mcmc_ocaml_b200 import Failure, InvalidArgument, nested, plugins as P

pytestmark = pytest.mark.gpu

PRIOR = P.box([0, 0], [1, 1], 0.0, closed=False)


@pytest.mark.parametrize("batch", [1, 16])
def test_point_for_point_vs_oracle(ctx, og, batch):
    like = P.gauss_diag([0.5, 0.5], [0.1, 0.1])
    ctx.set_seed(21)
    g = nested.nested_evidence(like, PRIOR, [0, 0], [1, 1], nlive=128, nmcmc=30, batch=batch, ctx=ctx)
    o = og.nested_evidence(21, 0, like, PRIOR, [0, 0], [1, 1], nlive=128, nmcmc=30, batch=batch)
    assert len(g.log_likelihood) == len(o["ll"])
    # the likelihood uses log(sigma) (CUDA libm vs glibc): coordinates exact, ll to 1e-13
    assert np.array_equal(g.points, o["pts"])
    np.testing.assert_allclose(g.log_likelihood, o["ll"], rtol=1e-13, atol=1e-13)
    assert g.log_evidence == pytest.approx(o["log_ev"], abs=1e-11)
    assert g.log_delta_evidence == pytest.approx(o["log_dev"], abs=1e-9)
    np.testing.assert_allclose(g.log_weights, o["logw"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("D,nlive,batch,nmcmc,kind", [(3, 300, 40, 25, "diag"), (16, 400, 70, 20, "shell"), (5, 256, 33, 30, "shell"),
                                                      (20, 300, 64, 12, "diag"), (4, 200, 1, 40, "shell")])
def test_point_for_point_more_shapes(ctx, og, D, nlive, batch, nmcmc, kind):
    """odd and even D, D == DMAX and D < DMAX, batches that are not a multiple of the warp, closed box prior"""
    like = P.gauss_diag(np.full(D, 0.5), np.linspace(0.05, 0.2, D)) if kind == "diag" else P.shell(np.full(D, 0.5), 0.3, 0.05)
    prior = P.box(np.zeros(D), np.ones(D), 0.0)
    ctx.set_seed(1000 + D)
    g = nested.nested_evidence(like, prior, np.zeros(D), np.ones(D), nlive=nlive, nmcmc=nmcmc, batch=batch, ctx=ctx)
    o = og.nested_evidence(1000 + D, 0, like, prior, np.zeros(D), np.ones(D), nlive=nlive, nmcmc=nmcmc, batch=batch)
    assert len(g.log_likelihood) == len(o["ll"])
    assert np.array_equal(g.points, o["pts"])
    np.testing.assert_allclose(g.log_likelihood, o["ll"], rtol=1e-13, atol=1e-13)
    assert g.log_evidence == pytest.approx(o["log_ev"], abs=1e-10)


def test_chunked_chain_walk(ctx, og, monkeypatch):
    """nmcmc walked in several chunks (proposal buffer smaller than nmcmc x K): chain state parked in global memory"""
    D = 6
    like = P.shell(np.full(D, 0.5), 0.3, 0.05)
    prior = P.box(np.zeros(D), np.ones(D), 0.0)
    ctx.set_seed(77)
    whole = nested.nested_evidence(like, prior, np.zeros(D), np.ones(D), nlive=200, nmcmc=50, batch=24, ctx=ctx)
    monkeypatch.setenv("MCMC_GPU_NEST_CHUNK", "7")
    ctx.set_seed(77)
    parts = nested.nested_evidence(like, prior, np.zeros(D), np.ones(D), nlive=200, nmcmc=50, batch=24, ctx=ctx)
    assert np.array_equal(whole.points, parts.points) and np.array_equal(whole.log_likelihood, parts.log_likelihood)
    assert whole.log_evidence == parts.log_evidence


def test_weights_deterministic_parity(ctx, og):
    """N5: evidence_error_and_weights given identical inputs, K = 1 and batched"""
    rng = np.random.default_rng(3)
    for nlive, batch, nret in [(100, 1, 5000), (1000, 1, 40000), (512, 64, 64 * 300), (1000, 7, 7 * 1234)]:
        ll = np.sort(rng.normal(-20, 8, nret + nlive))
        lev, ldev, lw = nested.evidence_error_and_weights(ll, nlive, batch, ctx=ctx)
        olev, oldev, olw = og.nested_weights(ll, nlive, batch)
        assert lev == pytest.approx(olev, abs=1e-12 * max(1.0, abs(olev)))
        assert ldev == pytest.approx(oldev, abs=1e-9)
        np.testing.assert_allclose(lw, olw, rtol=0, atol=1e-11)
        assert np.exp(lw).sum() == pytest.approx(1.0, abs=1e-8)          # nested_test.ml:66-85
    assert nested.log_total_error_estimate(-1.3, -4.0, 1000) == pytest.approx(og.nested_log_total_error(-1.3, -4.0, 1000), rel=1e-15)


def test_single_gaussian_reference_defaults(ctx):
    """nested_test.ml:23-39 with the reference defaults nlive=1000, nmcmc=1000, epsrel=0.01, mode_hopping_frac=0.1;
    batch = 1 (the reference's schedule) and batch = 64.  The reference's criterion -- |Z - 1| within twice the
    reported error -- uses an error estimate without the information H (nested.ml:148-150: quadrature error and
    1/sqrt(nlive) only), so it fails for a fair share of seeds even on the reference's own schedule
    (profiles/r02/r02_cfg4_batches.json).  Here: every run lies within 3 sqrt(H / nlive) of the truth (the scatter
    nested sampling has), and the reference's criterion holds for the majority of the seeds."""
    like = P.gauss_diag([0.5, 0.5], [0.1, 0.1])
    for batch, seeds in [(64, (31, 32, 33, 34)), (1, (35, 36))]:
        ok_ref = 0
        for seed in seeds:
            ctx.set_seed(seed)
            r = nested.nested_evidence(like, PRIOR, [0, 0], [1, 1], batch=batch, nmcmc=1000 if batch > 1 else 100, ctx=ctx)
            ev = math.exp(r.log_evidence)
            err = math.exp(nested.log_total_error_estimate(r.log_evidence, r.log_delta_evidence, 1000))
            w = np.exp(r.log_weights)
            H = float(np.sum(w * r.log_likelihood) - r.log_evidence)
            assert 1.0 < H < 2.5                                                     # analytic: log(1 / (2 pi e sigma^2)) = 1.77
            assert abs(r.log_evidence) <= 3.0 * math.sqrt(H / 1000.0) + 0.01 and err < 0.1
            ok_ref += abs(ev - 1.0) <= 2.0 * err
            assert w.sum() == pytest.approx(1.0, abs=1e-8)
            assert np.sum(w * r.points[:, 0]) == pytest.approx(0.5, abs=0.1)
            post = nested.posterior_samples(100, r, ctx=ctx)                           # nested_test.ml:87-105
            assert len(r.log_likelihood) > 100 and post[:, 0].mean() == pytest.approx(0.5, abs=0.05)
        assert ok_ref >= (len(seeds) + 1) // 2


def test_four_gaussians(ctx):  # nested_test.ml:41-64
    like = P.gauss_mix([[0.25, 0.25], [0.25, 0.75], [0.75, 0.25], [0.75, 0.75]], [0.05, 0.05])
    ctx.set_seed(33)
    r = nested.nested_evidence(like, PRIOR, [0, 0], [1, 1], batch=50, ctx=ctx)
    ev = math.exp(r.log_evidence)
    err = math.exp(nested.log_total_error_estimate(r.log_evidence, r.log_delta_evidence, 1000))
    w = np.exp(r.log_weights)
    H = float(np.sum(w * r.log_likelihood) - r.log_evidence)
    # within the scatter nested sampling has (see test_single_gaussian_reference_defaults); the reference's own
    # criterion (nested_test.ml:63-64) is 2x its H-free error estimate
    assert abs(r.log_evidence - math.log(4.0)) <= 3.0 * math.sqrt(H / 1000.0) + 0.01 and err < 0.5


def test_gaussian_shell_config4_small(ctx):
    """BASELINE.json config 4 at test size: 16-D Gaussian shell (r=2, w=0.1)
    on [-6,6]^16, analytic Z by radial quadrature"""
    from scipy import integrate, special
    D, r0, w = 16, 2.0, 0.1
    like = P.shell(np.zeros(D), r0, w)
    prior = P.box(np.full(D, -6.0), np.full(D, 6.0), -D * math.log(12.0))
    area = 2 * math.pi ** (D / 2) / special.gamma(D / 2)
    Z, _ = integrate.quad(lambda r: area * r ** (D - 1) * math.exp(-(r - r0) ** 2 / (2 * w * w)) / math.sqrt(2 * math.pi * w * w), 0, 6)
    logZ = math.log(Z) - D * math.log(12.0)
    ctx.set_seed(34)
    res = nested.nested_evidence(like, prior, np.full(D, -6.0), np.full(D, 6.0), nlive=2000, nmcmc=300, batch=256, ctx=ctx)
    err = nested.log_total_error_estimate(res.log_evidence, res.log_delta_evidence, 2000) - res.log_evidence
    # log-evidence within 5 relative-error units (nmcmc=300 DE steps leave some correlation)
    assert abs(res.log_evidence - logZ) < max(0.5, 5 * math.exp(err))


def test_errors(ctx):
    like = P.gauss_diag([0.5, 0.5], [0.1, 0.1])
    with pytest.raises(InvalidArgument):
        nested.nested_evidence(like, PRIOR, [0, 0], [1, 1], nlive=10, batch=10, ctx=ctx)
    with pytest.raises(Failure):                     # output arrays too small
        nested.nested_evidence(like, PRIOR, [0, 0], [1, 1], nlive=100, nmcmc=10, max_points=150, ctx=ctx)


def test_posterior_samples_match_oracle(ctx, og):
    """nested.ml:152-178 (weight_binary_search_index) on the same Philox stream"""
    rng = np.random.default_rng(2)
    lw = np.log(rng.dirichlet(np.ones(5000)))
    fake = nested.NestedOutput(0.0, 0.0, np.arange(5000, dtype=float)[:, None], np.zeros(5000), np.zeros(5000), lw, 100, 1)
    ctx.set_seed(44)
    got = nested.posterior_samples(2000, fake, ctx=ctx)[:, 0].astype(np.int64)
    want = og.nested_posterior_indices(44, 0, lw, 2000)
    assert np.array_equal(got, want)
    # drawn in proportion to the weights
    counts = np.bincount(got, minlength=5000)
    top = np.argsort(lw)[-50:]
    assert counts[top].sum() / 2000 == pytest.approx(np.exp(lw[top]).sum(), abs=0.03)


def test_nested_first_call_on_a_fresh_context(og):
    """The draw-ahead of the next batch runs on the context's second stream, which exists from mg_ctx_create on:
    nested_evidence as the very first call of a fresh context (no sampler call before it) gives the oracle's run."""
    from mcmc_ocaml_b200 import Context
    like = P.gauss_diag([0.5, 0.5], [0.1, 0.1])
    with Context(0, 321) as c:
        g = nested.nested_evidence(like, PRIOR, [0, 0], [1, 1], nlive=128, nmcmc=30, batch=16, ctx=c)
    o = og.nested_evidence(321, 0, like, PRIOR, [0, 0], [1, 1], nlive=128, nmcmc=30, batch=16)
    assert np.array_equal(g.points, o["pts"])
    assert g.log_evidence == pytest.approx(o["log_ev"], abs=1e-11)


def test_plain_load_chain_kernel_equals_pipelined(ctx, monkeypatch):
    """The chain kernel compiled at run time for user plugins (nested_kernel_dev.cuh: plain loads) and the pipelined
    kernel of the built-in plugins read the same pre-drawn proposals and do the same arithmetic: same run."""
    like = P.shell(np.full(5, 0.5), 0.3, 0.05)
    prior = P.box(np.zeros(5), np.ones(5), 0.0)
    ctx.set_seed(77)
    a = nested.nested_evidence(like, prior, np.zeros(5), np.ones(5), nlive=300, nmcmc=40, batch=33, ctx=ctx)
    monkeypatch.setenv("MCMC_GPU_NEST_SIMPLE", "1")
    ctx.set_seed(77)
    b = nested.nested_evidence(like, prior, np.zeros(5), np.ones(5), nlive=300, nmcmc=40, batch=33, ctx=ctx)
    assert np.array_equal(a.points, b.points) and np.array_equal(a.log_likelihood, b.log_likelihood)
    assert a.log_evidence == b.log_evidence and np.array_equal(a.log_weights, b.log_weights)


def test_observer_sees_every_retired_point_in_order(ctx):
    """?observer (nested.ml:123-125,136): called with each retired point, in retirement order"""
    import ctypes as C
    like = P.gauss_diag([0.5, 0.5], [0.1, 0.1])
    seen = []
    CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_double, C.c_double)

    def cb(user, value, dim, ll, lp):
        seen.append(([value[i] for i in range(dim)], ll, lp))
    fn = CB(cb)
    ctx.check(ctx.lib.mg_nested_set_observer(ctx.h, fn, None))
    try:
        ctx.set_seed(9)
        r = nested.nested_evidence(like, PRIOR, [0, 0], [1, 1], nlive=100, nmcmc=20, batch=7, ctx=ctx)
    finally:
        ctx.check(ctx.lib.mg_nested_set_observer(ctx.h, None, None))
    nret = len(r.log_likelihood) - 100
    assert len(seen) == nret and nret > 0
    assert np.array_equal(np.array([s[0] for s in seen]), r.points[:nret])
    assert np.array_equal(np.array([s[1] for s in seen]), r.log_likelihood[:nret])
