"""The numpy restatement of ellipse.ml (oracle/ellipse_np.py) against the reference's own tests
(test/ellipse_test.ml:60-84) and against invariants of the recursion (ellipse.ml:150-173)."""
import numpy as np
import pytest

from oracle import ellipse_np as E


def test_eigensystem_reconstruction():
    """ellipse_test.ml:60-71: evecs . diag evals . evecs^T reproduces 101 random symmetric 5 x 5 matrices"""
    rng = np.random.default_rng(1)
    for _ in range(101):
        m = rng.uniform(-1.0, 1.0, (5, 5))
        m = np.triu(m) + np.triu(m, 1).T
        w, z = E.eigensystem(m)
        assert np.all(np.diff(w) >= 0.0)
        np.testing.assert_allclose(z @ np.diag(w) @ z.T, m, atol=1e-12)


def test_enclose():
    """ellipse_test.ml:73-84: every point of 1000 uniform 5-D points lies inside enclosing_ellipse 2.0"""
    rng = np.random.default_rng(2)
    for _ in range(20):
        pts = rng.uniform(-1.0, 1.0, (1000, 5))
        ell = E.enclosing_ellipse(2.0, pts)
        r = E.elliptical_ranges(ell, pts)
        assert np.all(r < 1.0)
        # the farthest point sits at 1 / sf^(1/D) exactly by construction (rescale_ellipse, ellipse.ml:83-86)
        assert r.max() == pytest.approx(2.0 ** (-1.0 / 5.0), rel=1e-12)
        assert E.elliptical_range(ell, pts[17]) == r[17]


def test_center_and_sigma_follow_the_reference_fold():
    rng = np.random.default_rng(3)
    pts = rng.normal(0.3, 2.0, (257, 3))
    cen, nf = np.zeros(3), 257.0
    for p in pts:                                     # ellipse.ml:41-44
        cen = cen + p / nf
    assert np.array_equal(E.center(pts), cen)
    s = np.zeros((3, 3))
    for p in pts:                                     # ellipse.ml:56-64
        for j in range(3):
            dxj = p[j] - cen[j]
            s[j, j] += dxj * dxj / nf
            for k in range(j + 1, 3):
                d = dxj * (p[k] - cen[k]) / nf
                s[j, k] += d; s[k, j] += d
    assert np.array_equal(E.sigma2(cen, pts), s)


def test_tree_invariants_and_circumcircles():
    rng = np.random.default_rng(4)
    pts = rng.normal(0.0, 1.0, (3000, 4)) * np.array([1.0, 2.0, 0.5, 3.0])
    t = E.ellipse_tree(1.5, pts)
    flat = E.flatten(t)
    n = len(flat["ids"])
    assert np.array_equal(flat["ids"][0], np.arange(3000))
    for k in range(n):
        ids = flat["ids"][k]
        assert len(ids) >= 5 and np.all(np.diff(ids) > 0)            # >= ndim + 1 points, input order kept
        ell = E.Ellipse(flat["center"][k], flat["axes"][k], flat["orientation"][k])
        assert np.all(E.elliptical_ranges(ell, pts[ids]) < 1.0)
        split = E.widest_dimension(ell)
        assert split == 3                                             # ascending eigenvalues: always the last index
        l, r = flat["left"][k], flat["right"][k]
        lo = ids[pts[ids, split] < ell.center[split]]
        hi = ids[~(pts[ids, split] < ell.center[split])]
        assert (l < 0 and len(lo) < 5) or np.array_equal(flat["ids"][l], lo)
        assert (r < 0 and len(hi) < 5) or np.array_equal(flat["ids"][r], hi)
        if l < 0 and r < 0:
            assert flat["cc_radius"][k] == ell.axes.max() and np.array_equal(flat["cc_center"][k], ell.center)
    # union_circumcircles (ellipse.ml:112-127): containment cases and the general formula as coded
    c1, c2 = np.zeros(2), np.array([1.0, 0.0])
    assert E.union_circumcircles((c1, 5.0), (c2, 1.0))[1] == 5.0
    assert E.union_circumcircles((c1, 1.0), (c2, 5.0))[1] == 5.0
    c, r = E.union_circumcircles((c1, 1.0), (c2, 1.0))
    assert r == 3.0 and np.array_equal(c, [0.5, 0.0])
    assert E.in_circumcircle([0.4, 0.0], (c1, 0.5)) and not E.in_circumcircle([0.5, 0.0], (c1, 0.5))


def test_unsplittable_node_is_reported():
    pts = np.ones((20, 2))
    with pytest.raises(RecursionError):
        E.ellipse_tree(2.0, pts)
