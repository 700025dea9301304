"""The oracle's kd_tree.ml / interpolate_pdf.ml restatement against the
reference's own structural tests (test/kd_tree_test.ml, interpolate_pdf_test.ml)
and the exact split rule of SURVEY.md F7."""
import numpy as np
import pytest


def children_boxes(t, ex, node, lo, hi):
    d, s = ex["split_dim"][node], ex["split_val"][node]
    lhi = hi.copy(); lhi[d] = s
    rlo = lo.copy(); rlo[d] = s
    return (lo, lhi), (rlo, hi)


def check_invariant(og, pts):
    """test/kd_tree_test.ml:45-64 acceptable_tree_p: every object of a split
    cell lies in exactly one child's box."""
    lo, hi = pts.min(0), pts.max(0)
    t = og.Tree(pts, lo, hi)
    ex = t.export()
    stack = [(0, lo.copy(), hi.copy())]
    while stack:
        node, l, h = stack.pop()
        if ex["left"][node] < 0:
            continue
        (ll, lh), (rl, rh) = children_boxes(t, ex, node, l, h)
        objs = pts[ex["perm"][ex["begin"][node]:ex["end"][node]]]
        in_l = np.all((objs >= ll) & (objs <= lh), axis=1)
        in_r = np.all((objs >= rl) & (objs <= rh), axis=1)
        assert np.all(in_l ^ in_r)
        L = ex["left"][node]
        # children partition the parent's range, left-heavy: floor(n/2)+1 / rest (F7) for distinct keys
        n = ex["end"][node] - ex["begin"][node]
        nl = ex["end"][L] - ex["begin"][L]
        assert ex["begin"][L] == ex["begin"][node] and ex["end"][L] == ex["begin"][L + 1]
        assert ex["end"][L + 1] == ex["end"][node]
        assert nl == (n // 2 + 1 if n > 2 else 1)
        stack.append((L, ll, lh)); stack.append((L + 1, rl, rh))
    return t


def test_tree_invariant(og):
    rng = np.random.default_rng(0)
    for _ in range(20):
        check_invariant(og, rng.random((250, 2)))


def test_tree_depth(og):  # kd_tree_test.ml:73-78
    pts = np.random.default_rng(1).random((1024, 2))
    t = og.Tree(pts, pts.min(0), pts.max(0))
    assert 8 <= t.info()["nlevels"] <= 12


def test_split_rule_small_cases(og):
    """F7: n=4 -> 3/1, n=3 -> 2/1, n=2 -> 1/1 through adjust_for_empty_split"""
    for n, want in [(4, (3, 1)), (3, (2, 1)), (2, (1, 1))]:
        pts = np.arange(n, dtype=float).reshape(-1, 1)
        ex = og.Tree(pts, [0.0], [float(n)]).export()
        L = ex["left"][0]
        assert (ex["end"][L] - ex["begin"][L], ex["end"][L + 1] - ex["begin"][L + 1]) == want
    # split value = 0.5 * (max left + min right), kd_tree.ml:113
    ex = og.Tree(np.array([[0.0], [1.0], [4.0], [9.0]]), [0.0], [9.0]).export()
    assert ex["split_val"][0] == 0.5 * (4.0 + 9.0)


def test_duplicates_make_leaves(og):
    pts = np.array([[0.5, 0.5]] * 5 + [[0.25, 0.75]] * 3)
    t = og.Tree(pts, [0, 0], [1, 1])
    ex = t.export()
    leaves = np.where(ex["left"] < 0)[0]
    assert sorted(ex["end"][leaves] - ex["begin"][leaves]) == [3, 5]     # kd_tree.ml:159-160
    # jump_prob counts the duplicates: nobjs / (V N)
    p = t.jump_prob(np.array([[0.5, 0.5]]))
    node, lo, hi = t.find_cell(np.array([[0.5, 0.5]]), boxes=True)
    assert p[0] == 5 / (og.bounds_volume(lo[0], hi[0]) * 8)


def test_interp_draw_linear_pdf(og):  # interpolate_pdf_test.ml:43-53
    rng = np.random.default_rng(5)
    b0, b1 = 1.3, 2.9
    pts = np.stack([np.sqrt(b0 * b0 * rng.random(10000)), np.sqrt(b1 * b1 * rng.random(10000))], axis=1)
    t = og.Tree(pts, [0.0, 0.0], [b0, b1])
    d = t.draw(11, 0, 10000)
    assert d[:, 0].mean() == pytest.approx(2 / 3 * b0, rel=0.05)
    assert d[:, 1].mean() == pytest.approx(2 / 3 * b1, rel=0.05)


def test_jump_prob_integrates_to_one(og):
    rng = np.random.default_rng(6)
    pts = rng.normal(0.5, 0.1, (2000, 2)).clip(0.01, 0.99)
    t = og.Tree(pts, [0, 0], [1, 1])
    q = rng.random((200000, 2))
    assert t.jump_prob(q).mean() == pytest.approx(1.0, rel=0.05)     # Monte Carlo integral over the unit box
    assert t.jump_prob(q, nstop=64).mean() == pytest.approx(1.0, rel=0.05)
