"""The fast C++ oracle (liboracle.so) against a second, independent, deliberately naive list-based transcription of
kd_tree.ml / interpolate_pdf.ml / evidence.ml (oracle/naive_ml.py).  The OCaml reference cannot be run here, so the
restatement that every GPU parity test leans on is itself cross-checked by a transcription that shares no code and
no data structure with it: EXACT equality of the flat trees, the object order, cell assignment, densities and the
three evidence estimators."""
import math

import numpy as np
import pytest

from oracle import naive_ml as nv


def _cases():
    rng = np.random.default_rng(2024)
    out = []
    out.append(("uniform 2-D", rng.random((700, 2))))
    out.append(("gaussian 5-D", rng.normal(0.5, 0.1, (1500, 5))))
    out.append(("grid ties", np.round(rng.random((600, 2)), 1)))
    x = rng.normal(0.0, 1.0, (900, 3))
    rep = rng.random(900) < 0.35
    rep[0] = False
    idx = np.arange(900); idx[rep] = 0; idx = np.maximum.accumulate(idx)
    out.append(("MH repeats 3-D", x[idx]))
    out.append(("signed zeros", np.array([[0.0, -0.0], [-0.0, 0.0], [0.0, 1.0], [1.0, -0.0], [0.5, 0.5]])))
    out.append(("two clusters", np.concatenate([np.zeros((20, 2)), np.ones((20, 2))])))
    out.append(("1-D descending", np.sort(rng.random((257, 1)), axis=0)[::-1].copy()))
    out.append(("20-D", rng.normal(0.5, 0.05, (800, 20))))
    return out


@pytest.mark.parametrize("name,pts", _cases(), ids=[c[0] for c in _cases()])
def test_tree_flat_arrays_identical(og, name, pts):
    n, d = pts.shape
    lo, hi = (pts.min(0) - 0.25).tolist(), (pts.max(0) + 0.25).tolist()
    objs = [(i, pts[i].tolist()) for i in range(n)]
    kd = nv.KdTree(lambda o: o[1])
    tree = kd.tree_of_objects(objs, lo, hi)
    flat = nv.flatten_bfs(tree, lambda o: o[0])
    ex = og.Tree(pts, lo, hi).export()
    assert np.array_equal(ex["split_dim"], np.array(flat["split_dim"], np.int32))
    assert np.array_equal(ex["split_val"], np.array(flat["split_val"]))          # bit-exact (0.5 * (a + b))
    assert np.array_equal(ex["left"], np.array(flat["left"], np.int32))
    assert np.array_equal(ex["begin"], np.array(flat["begin"], np.int32))
    assert np.array_equal(ex["end"], np.array(flat["end"], np.int32))
    assert np.array_equal(ex["perm"], np.array(flat["perm"], np.int32))          # List.partition order


@pytest.mark.parametrize("d", [1, 2, 6])
def test_find_cell_and_jump_prob_identical(og, d):
    rng = np.random.default_rng(d)
    pts = rng.random((1200, d))
    pts[100:130] = pts[100]                         # a leaf of duplicates
    lo, hi = [0.0] * d, [1.0] * d
    ip = nv.InterpolatePdf([p.tolist() for p in pts], lo, hi)
    o = og.Tree(pts, lo, hi)
    q = np.concatenate([rng.random((400, d)), pts[:200], rng.normal(0.5, 0.8, (100, d))])
    want = np.array([ip.jump_prob(p.tolist()) for p in q])
    assert np.array_equal(o.jump_prob(q), want)
    # the located cell, identified by its object list
    ex = o.export()
    node = o.find_cell(q)
    for k in range(0, len(q), 7):
        objs = ip.find_cell(q[k].tolist())[0]
        got = pts[ex["perm"][ex["begin"][node[k]]:ex["end"][node[k]]]]
        assert np.array_equal(np.array(objs), got)
    for nmax in (1, 16, 100):
        ok = o.find_cell(q, nmax) >= 0              # a leaf of > nmax duplicates: the reference raises
        for k in np.nonzero(~ok)[0][:3]:
            with pytest.raises(RuntimeError):
                ip.jump_prob_high_level(nmax, q[k].tolist())
        want = np.array([ip.jump_prob_high_level(nmax, p.tolist()) for p in q[ok]])
        assert np.array_equal(o.jump_prob(q[ok], nmax), want)


def _samples(rng, n, d, repeat):
    x = rng.normal(0.5, 0.1, (n, d))
    if repeat:
        rep = rng.random(n) < repeat
        rep[0] = False
        idx = np.arange(n); idx[rep] = 0; idx = np.maximum.accumulate(idx)
        x = x[idx]
    ll = (-0.91893853320467274178 - math.log(0.1) - 0.5 * ((x - 0.5) / 0.1) ** 2).sum(1)
    lp = -0.3 * np.abs(x).sum(1)
    return x, ll, lp


@pytest.mark.parametrize("n,d,repeat,nmax,eps", [(3000, 2, 0.0, 64, 0.1), (2500, 3, 0.4, 16, 0.1), (2000, 5, 0.2, 64, 1e-3),
                                                 (500, 1, 0.0, 8, 0.1), (4000, 2, 0.3, 64, 1e9)])
def test_evidence_estimators_identical(og, n, d, repeat, nmax, eps):
    x, ll, lp = _samples(np.random.default_rng(n + d), n, d, repeat)
    samples = [(x[i].tolist(), float(ll[i]), float(lp[i])) for i in range(n)]
    ev = nv.Evidence()
    # harmonic mean: the reference's left-to-right fold
    assert og.evidence_harmonic_mean(ll)[0] == ev.evidence_harmonic_mean(samples)
    # direct
    o = og.evidence_direct(x, ll, lp, n=nmax)
    assert o["value"] == ev.evidence_direct(samples, n=nmax)
    # (the oracle's full_tree diagnostic re-partitions perm below the cells, so its per-cell fold order is not the
    # reference's list order: equal to rounding only)
    assert og.evidence_direct(x, ll, lp, n=nmax, full_tree=True)["value"] == pytest.approx(o["value"], rel=1e-14)
    # Lebesgue (Weinberg)
    want, nkept, ncells = ev.evidence_lebesgue(samples, n=nmax, eps=eps)
    o = og.evidence_lebesgue(x, ll, lp, n=nmax, eps=eps)
    assert (o["value"], o["nkept"], o["ncells"]) == (want, nkept, ncells)
    assert og.evidence_lebesgue(x, ll, lp, n=nmax, eps=eps, full_tree=True)["value"] == want


def test_bounds_volume_and_reference_structure(og):
    """kd_tree_test.ml:45-78 on the naive tree itself (sanity of the transcription)"""
    rng = np.random.default_rng(5)
    pts = rng.random((1024, 2))
    kd = nv.KdTree(lambda o: o)
    tree = kd.tree_of_objects([p.tolist() for p in pts], [0.0, 0.0], [1.0, 1.0])

    def depth(t):
        return 0 if t is None else 1 + max(depth(t[3]), depth(t[4]))

    assert 8 <= depth(tree) - 1 <= 12 or 8 <= depth(tree) <= 12

    def check(t):
        if t is None or (t[3] is None and t[4] is None):
            return
        for o in t[0]:
            in_l = nv.InterpolatePdf.in_bounds(o, t[3][1], t[3][2])
            in_r = nv.InterpolatePdf.in_bounds(o, t[4][1], t[4][2])
            assert in_l != in_r
        check(t[3]); check(t[4])

    check(tree)
    assert kd.bounds_volume([0.0, -1.0, 0.5], [1.0, 2.0, 0.75]) == og.bounds_volume([0.0, -1.0, 0.5], [1.0, 2.0, 0.75])


@pytest.mark.parametrize("n,nlive,seed", [(60, 10, 1), (5000, 1000, 2), (2001, 2000, 3), (350, 7, 4)])
def test_nested_weights_identical(og, n, nlive, seed):
    """nested.ml:81-120 / 148-150 transcribed statement by statement in Python against oracle.cpp (the reference-order
    scan, K = 1): log evidence, its error and every log weight bit for bit."""
    rng = np.random.default_rng(seed)
    ll = np.sort(rng.normal(-20.0, 8.0, n))
    ll[n // 3] = ll[n // 3 + 1]                                   # a tie
    lev, ldev, lw = og.nested_weights(ll, nlive)
    nev, ndev, nw = nv.evidence_error_and_weights(nlive, [float(v) for v in ll])
    assert lev == nev and ldev == ndev
    assert np.array_equal(lw, np.array(nw))
    assert og.nested_log_total_error(lev, ldev, nlive) == nv.log_total_error_estimate(nev, ndev, nlive)
    # weight_binary_search_index (nested.ml:152-165) on the running sums of the weights
    sums = np.cumsum(np.exp(lw))
    for x in (0.0, float(sums[0]), 0.3, 0.999999, 1.5):
        i = nv.weight_binary_search_index(x, [float(v) for v in sums])
        assert 0 <= i < n and (i == 0 or sums[i - 1] < x or i == n - 1) and (x <= sums[i] or i == n - 1)
