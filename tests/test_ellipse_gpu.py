"""Ellipse on the GPU (mg_ellipse_*, csrc/ellipse.cu) against the numpy restatement of ellipse.ml and against the
reference's own tests (test/ellipse_test.ml:60-84).

Tolerances (stated, float64): the GPU sums per slice and divides once, the reference divides every term and folds
left to right (ellipse.ml:41-44,:56-64) -> centre and covariance agree to 1e-13 relative; the eigen-system comes from
cyclic Jacobi here and from LAPACK in the oracle (the reference: LAPACK dsyevr through Lacaml) -> eigenvalues /
axes to 1e-11 relative, elliptical ranges to 1e-9 relative; the tree's STRUCTURE (children, point sets) is exact."""
import numpy as np
import pytest

from mcmc_ocaml_b200 import Failure, InvalidArgument, ellipse
from oracle import ellipse_np as E

pytestmark = pytest.mark.gpu


def test_reference_enclose_test(ctx):
    """ellipse_test.ml:73-84 on the GPU: 101 clouds of 1000 uniform 5-D points, every point inside (r < 1)"""
    rng = np.random.default_rng(5)
    for _ in range(101):
        pts = rng.uniform(-1.0, 1.0, (1000, 5))
        ell = ellipse.enclosing_ellipse(2.0, pts, ctx=ctx)
        r = ellipse.elliptical_range(ell, pts, ctx=ctx)
        assert np.all(r < 1.0)
        assert r.max() == pytest.approx(2.0 ** (-1.0 / 5.0), rel=1e-12)


def test_reference_eigensystem_test(ctx):
    """ellipse_test.ml:60-71 through enclosing_ellipse: orientation . diag(eigenvalues) . orientation^T reproduces the
    covariance, the orientation is orthonormal, eigenvalues ascend"""
    rng = np.random.default_rng(6)
    for d in (1, 2, 3, 5, 8, 16, 32):
        a = rng.normal(size=(d, d))
        pts = rng.normal(size=(4000, d)) @ a + rng.normal(size=d)
        ell = ellipse.enclosing_ellipse(1.0, pts, ctx=ctx)
        cen = E.center(pts)
        sig = E.sigma2(cen, pts)
        np.testing.assert_allclose(ell.center, cen, rtol=1e-13, atol=1e-13)
        scale = ell.axes[-1] / np.linalg.eigvalsh(sig)[-1]                    # axes = eigenvalues * sf^(1/D) * r_max
        w = ell.axes / scale
        assert np.all(np.diff(w) >= 0.0)
        np.testing.assert_allclose(ell.orientation @ np.diag(w) @ ell.orientation.T, sig, rtol=0, atol=1e-12 * np.abs(sig).max())
        np.testing.assert_allclose(ell.orientation.T @ ell.orientation, np.eye(d), atol=1e-13)


@pytest.mark.parametrize("n,d,sf", [(1000, 5, 2.0), (50_000, 3, 1.5), (20_000, 10, 2.0), (7, 2, 3.0), (4097, 17, 1.1)])
def test_enclosing_ellipse_matches_oracle(ctx, n, d, sf):
    rng = np.random.default_rng(n + d)
    pts = rng.normal(size=(n, d)) * rng.uniform(0.2, 3.0, d) + rng.uniform(-5, 5, d)
    g = ellipse.enclosing_ellipse(sf, pts, ctx=ctx)
    o = E.enclosing_ellipse(sf, pts)
    np.testing.assert_allclose(g.center, o.center, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(g.axes, o.axes, rtol=1e-11)
    q = rng.normal(size=(500, d)) * 3.0
    np.testing.assert_allclose(ellipse.elliptical_range(g, q, ctx=ctx), E.elliptical_ranges(o, q), rtol=1e-9)
    # the range kernel itself, on the oracle's ellipse: same operation order -> same bits
    assert np.array_equal(ellipse.elliptical_range(ellipse.Ellipse(o.center, o.axes, o.orientation), q, ctx=ctx),
                          E.elliptical_ranges(o, q))
    assert ellipse.elliptical_range(g, q[3], ctx=ctx) == ellipse.elliptical_range(g, q, ctx=ctx)[3]


@pytest.mark.parametrize("n,d,sf,seed", [(3000, 4, 1.5, 1), (20_000, 5, 2.0, 2), (100_000, 3, 2.0, 3), (5000, 12, 1.2, 4), (64, 2, 2.0, 5)])
def test_ellipse_tree_matches_oracle(ctx, n, d, sf, seed):
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.normal(size=(n // 2, d)) * rng.uniform(0.3, 2.0, d),
                          rng.normal(size=(n - n // 2, d)) * 0.4 + rng.uniform(-3, 3, d)])     # two clusters
    t = ellipse.ellipse_tree(sf, pts, ctx=ctx)
    a = t.export()
    o = E.flatten(E.ellipse_tree(sf, pts))
    assert t.nnodes == len(o["ids"]) and t.npoints == n and t.dim == d
    assert np.array_equal(a["left"], o["left"]) and np.array_equal(a["right"], o["right"])     # structure: exact
    for k in range(t.nnodes):
        assert np.array_equal(t.points_of(k), o["ids"][k])                                      # point sets: exact
    assert np.array_equal(np.sort(a["perm"]), np.arange(n))
    np.testing.assert_allclose(a["center"], o["center"], rtol=1e-12, atol=1e-12)
    # axes = eigenvalue * sf^(1/D) * r_max, and r_max divides by the SMALLEST eigenvalue, which LAPACK (the oracle's
    # and the reference's solver) only knows to eps * |Sigma|: nodes of D + 1 .. 2 D points have condition numbers up
    # to 1e8, so the agreed tolerance scales with the node's condition number
    cond = o["axes"].max(axis=1) / o["axes"].min(axis=1)
    rel = np.abs(a["axes"] - o["axes"]).max(axis=1) / o["axes"].max(axis=1)
    assert np.all(rel <= 1e-12 * cond + 1e-11), float((rel / cond).max())
    assert np.all(np.abs(a["cc_radius"] - o["cc_radius"]) <= (1e-11 * cond.max() + 1e-10) * o["cc_radius"])
    well = cond < 1e3                                                                            # well-conditioned nodes
    np.testing.assert_allclose(a["axes"][well], o["axes"][well], rtol=1e-10)
    np.testing.assert_allclose(a["cc_center"], o["cc_center"], rtol=1e-7, atol=1e-8 * cond.max())
    for k in (0, t.nnodes // 2, t.nnodes - 1):                                                  # every point inside its node's ellipse
        ids = t.points_of(k)
        assert np.all(ellipse.elliptical_range(t.ellipse_of(k), pts[ids], ctx=ctx) < 1.0)
    t.close()


def test_ellipse_errors(ctx):
    with pytest.raises(InvalidArgument):            # assert (Array.length pts >= ndim + 1), ellipse.ml:152
        ellipse.ellipse_tree(2.0, np.zeros((3, 3)) + np.arange(3), ctx=ctx)
    with pytest.raises(Failure):                    # all points on one side of the centre: the reference never returns
        ellipse.ellipse_tree(2.0, np.ones((20, 2)), ctx=ctx)
    with pytest.raises(InvalidArgument):
        ellipse.enclosing_ellipse(2.0, np.zeros((10, 33)), ctx=ctx)
    # the context stays usable
    ell = ellipse.enclosing_ellipse(2.0, np.random.default_rng(0).normal(size=(100, 2)), ctx=ctx)
    assert np.all(ell.axes > 0)
