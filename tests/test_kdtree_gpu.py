"""GPU kd-tree build and Interpolate_pdf kernels against the oracle, through
the C ABI: flat node arrays, object order and cell assignment bit-exact;
densities and draws bit-exact (only + - * / are involved)."""
import numpy as np
import pytest

from mcmc_ocaml_b200 import InvalidArgument, interpolate_pdf, kd_tree

pytestmark = pytest.mark.gpu

KEYS = ["split_dim", "split_val", "left", "begin", "end", "perm"]


def assert_same_tree(gpu_tree, oracle_tree):
    a, b = gpu_tree.export(), oracle_tree.export()
    assert gpu_tree.nnodes == oracle_tree.info()["nnodes"]
    assert gpu_tree.nlevels == oracle_tree.info()["nlevels"]
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k


def mh_like(rng, n, d, repeat=0.3):
    """samples with runs of exact repeats, like a Metropolis-Hastings chain"""
    x = rng.normal(0.5, 0.1, (n, d))
    rep = rng.random(n) < repeat
    rep[0] = False
    idx = np.arange(n)
    idx[rep] = 0
    idx = np.maximum.accumulate(idx)
    return x[idx]


@pytest.mark.parametrize("n,d", [(1, 2), (2, 1), (3, 3), (4, 2), (17, 1), (250, 2), (1024, 2), (5000, 7),
                                 (20000, 20), (3001, 33), (2000, 64)])
def test_build_bit_exact_random(ctx, og, n, d):
    pts = np.random.default_rng(n * 131 + d).random((n, d))
    lo, hi = np.zeros(d), np.ones(d)
    assert_same_tree(kd_tree.KdTree(pts, lo, hi, ctx=ctx), og.Tree(pts, lo, hi))


def test_build_bit_exact_ties_and_duplicates(ctx, og):
    rng = np.random.default_rng(3)
    cases = [
        mh_like(rng, 4000, 3, 0.5),                               # exact repeated rows
        np.round(rng.random((3000, 2)), 1),                       # heavy ties on a 11x11 grid
        np.repeat(np.array([[0.25, 0.75, 0.5]]), 40, axis=0),     # all identical -> single leaf
        np.concatenate([np.zeros((50, 2)), np.ones((50, 2))]),    # two clusters of identical points
        np.stack([np.zeros(300), rng.random(300)], axis=1),       # one degenerate dimension
        np.array([[0.0, -0.0], [-0.0, 0.0], [0.0, 1.0], [1.0, -0.0]]),   # signed zeros compare equal
        np.sort(rng.random((999, 1)), axis=0)[::-1].copy(),       # descending input
    ]
    for pts in cases:
        d = pts.shape[1]
        lo, hi = pts.min(0) - 0.5, pts.max(0) + 0.5
        assert_same_tree(kd_tree.KdTree(pts, lo, hi, ctx=ctx), og.Tree(pts, lo, hi))


@pytest.mark.parametrize("min_split", [2, 8, 64, 1000])
def test_truncated_build(ctx, og, min_split):
    """Evidence never looks below the first cell with < n objects (evidence.ml:83-89)"""
    pts = mh_like(np.random.default_rng(9), 6000, 4)
    lo, hi = pts.min(0), pts.max(0)
    g, o = kd_tree.KdTree(pts, lo, hi, min_split=min_split, ctx=ctx), og.Tree(pts, lo, hi, min_split=min_split)
    assert_same_tree(g, o)
    ex = g.export()
    leaves = ex["left"] < 0
    cnt = ex["end"] - ex["begin"]
    assert cnt[leaves].sum() == len(pts) and np.array_equal(np.sort(ex["perm"]), np.arange(len(pts)))


def test_large_build(ctx, og):
    pts = np.random.default_rng(77).normal(0.5, 0.05, (300000, 5))
    lo, hi = np.zeros(5), np.ones(5)
    assert_same_tree(kd_tree.KdTree(pts, lo, hi, ctx=ctx), og.Tree(pts, lo, hi))


def test_reference_structural_tests(ctx):
    """test/kd_tree_test.ml:45-78 on the GPU tree"""
    rng = np.random.default_rng(1)
    pts = rng.random((1024, 2))
    t = kd_tree.KdTree(pts, pts.min(0), pts.max(0), ctx=ctx)
    assert 8 <= t.depth() <= 12
    ex = t.export()
    lo0, hi0 = pts.min(0), pts.max(0)
    stack = [(0, lo0.copy(), hi0.copy())]
    while stack:
        node, l, h = stack.pop()
        if ex["left"][node] < 0:
            continue
        d, s = ex["split_dim"][node], ex["split_val"][node]
        lh = h.copy(); lh[d] = s
        rl = l.copy(); rl[d] = s
        objs = pts[ex["perm"][ex["begin"][node]:ex["end"][node]]]
        in_l = np.all((objs >= l) & (objs <= lh), axis=1)
        in_r = np.all((objs >= rl) & (objs <= h), axis=1)
        assert np.all(in_l ^ in_r)
        stack += [(ex["left"][node], l, lh), (ex["left"][node] + 1, rl, h)]


def test_volume(ctx, og):
    pts = np.random.default_rng(2).random((100, 3))
    t = kd_tree.KdTree(pts, [0.0, -1.0, 0.5], [1.0, 2.0, 0.75], ctx=ctx)
    assert t.volume() == og.bounds_volume([0.0, -1.0, 0.5], [1.0, 2.0, 0.75]) == 1.0 * 3.0 * 0.25
    assert kd_tree.bounds_volume([0, 0], [0.1, 0.3]) == og.bounds_volume([0, 0], [0.1, 0.3])


def test_invalid_input(ctx):
    with pytest.raises(InvalidArgument):
        kd_tree.KdTree(np.array([[0.1, np.nan]]), [0, 0], [1, 1], ctx=ctx)
    with pytest.raises(InvalidArgument):
        kd_tree.KdTree(np.zeros((0, 2)), [0, 0], [1, 1], ctx=ctx)       # tree_of_objects []


@pytest.mark.parametrize("d", [1, 2, 5, 20, 40])
def test_find_cell_and_jump_prob_bit_exact(ctx, og, d):
    rng = np.random.default_rng(d)
    pts = mh_like(rng, 5000, d, 0.2).clip(0.0, 1.0)
    lo, hi = np.zeros(d), np.ones(d)
    g = interpolate_pdf.InterpPdf(pts, lo, hi, ctx=ctx)
    o = og.Tree(pts, lo, hi)
    q = np.concatenate([rng.random((3000, d)),                 # in the prior box
                        pts[:500],                             # the stored samples themselves
                        rng.normal(0.5, 1.0, (500, d)),        # partly outside the root box
                        pts[:200] + 0.0])
    assert np.array_equal(g.find_cell(q), o.find_cell(q))
    assert np.array_equal(g.jump_prob(q), o.jump_prob(q))
    from mcmc_ocaml_b200 import Failure
    for n in (1, 10, 64, 100000):
        want = o.find_cell(q, n)
        if np.any(want < 0):
            # a leaf of duplicates with more than n objects: the reference raises
            # Failure "encountered empty tree!" (interpolate_pdf.ml:124,149)
            with pytest.raises(Failure):
                g.find_cell(q, n)
            ok = want >= 0
            assert np.array_equal(g.find_cell(q[ok], n), want[ok])
            assert np.array_equal(g.jump_prob_high_level(n, q[ok]), o.jump_prob(q[ok], n))
        else:
            assert np.array_equal(g.find_cell(q, n), want)
            assert np.array_equal(g.jump_prob_high_level(n, q), o.jump_prob(q, n))


def test_split_plane_rounding_goes_left(ctx, og):
    """when 0.5*(a+b) rounds onto b, a point stored in the right child is
    located in the LEFT child (find_cell tests the left box inclusively,
    interpolate_pdf.ml:96-99,106; SURVEY 'hard parts')"""
    a = 1.0
    b = np.nextafter(a, 2.0)
    pts = np.array([[a], [b]])
    g = interpolate_pdf.InterpPdf(pts, [0.0], [2.0], ctx=ctx)
    o = og.Tree(pts, [0.0], [2.0])
    assert np.array_equal(g.find_cell(pts), o.find_cell(pts))
    assert np.array_equal(g.jump_prob(pts), o.jump_prob(pts))


def test_draw_bit_exact_and_distribution(ctx, og):
    rng = np.random.default_rng(5)
    b0, b1 = 1.3, 2.9
    pts = np.stack([np.sqrt(b0 * b0 * rng.random(10000)), np.sqrt(b1 * b1 * rng.random(10000))], axis=1)
    g = interpolate_pdf.InterpPdf(pts, [0.0, 0.0], [b0, b1], ctx=ctx)
    o = og.Tree(pts, [0.0, 0.0], [b0, b1])
    ctx.set_seed(31)
    d = g.draw(10000)
    assert np.array_equal(d, o.draw(31, 0, 10000))
    # interpolate_pdf_test.ml:43-53: mean of the linear pdf to 5 %
    assert d[:, 0].mean() == pytest.approx(2 / 3 * b0, rel=0.05)
    assert d[:, 1].mean() == pytest.approx(2 / 3 * b1, rel=0.05)
    ctx.set_seed(32)
    assert np.array_equal(g.draw_high_level(64, 5000), o.draw(32, 0, 5000, nstop=64))


def test_blob_round_trip(ctx, og):
    """a tree serialised to one device blob (what NCCL broadcasts) and rebuilt"""
    pts = np.random.default_rng(8).random((3000, 3))
    g = kd_tree.KdTree(pts, np.zeros(3), np.ones(3), ctx=ctx)
    p, n = g.blob()
    g2 = kd_tree.KdTree.from_blob(p, n, ctx=ctx)
    a, b = g.export(), g2.export()
    for k in KEYS:
        assert np.array_equal(a[k], b[k])
    q = np.random.default_rng(9).random((100, 3))
    i1 = interpolate_pdf.InterpPdf(None, None, None, tree=g)
    i2 = interpolate_pdf.InterpPdf(None, None, None, tree=g2)
    assert np.array_equal(i1.jump_prob(q), i2.jump_prob(q))


@pytest.mark.parametrize("cluster,what", [(40, "tie runs fixed up"), (300, "long run: 64-bit fallback"), (0, "MH repeats"),
                                          (-1, "both signs, zeros, denormals")])
def test_window_sort_ties(ctx, og, cluster, what):
    """Builds of >= 65,536 points sort each coordinate on a 32-bit window of the keys and repair runs of equal
    windows with the full keys (csrc/kdtree.cu tie_fix_kernel).  Wide-range data (windows truncate low mantissa
    bits) with clusters of points that differ only BELOW the window, in scrambled order, plus exact duplicates:
    the tree must still be the oracle's, object order included."""
    rng = np.random.default_rng(4242 + cluster)
    n, d = 70000, 3
    if cluster == -1:
        pts = rng.normal(0.0, 1.0, (n, d)) * np.array([1.0, 1e-300, 1e5])
        pts[::97, 0] = 0.0
        pts[::101, 0] = -0.0
        pts[::89, 1] = 5e-324 * rng.integers(0, 50, len(pts[::89]))
    elif cluster == 0:
        pts = mh_like(rng, n, d, repeat=0.4)
        pts[:, 1] = pts[:, 1] * 1e3          # a second dimension with a wide exponent range
    else:
        pts = rng.random((n, d)) * np.array([1.0, 100.0, 1e-3])
        for c in range(60):                  # clusters: same window, different low bits, shuffled
            at = rng.integers(0, n - cluster)
            centre = rng.random(d) * np.array([1.0, 100.0, 1e-3])
            k = rng.permutation(cluster)[:, None] * np.array([2.0 ** -50, 2.0 ** -44, 2.0 ** -60])
            pts[at:at + cluster] = centre * (1.0 + 0.0) + k * np.array([1.0, 1.0, 1.0])
        pts[1000:1010] = pts[1000]           # exact duplicates inside the data
    lo = pts.min(0) - 1.0
    hi = pts.max(0) + 1.0
    assert_same_tree(kd_tree.KdTree(pts, lo, hi, ctx=ctx), og.Tree(pts, lo, hi))


@pytest.mark.parametrize("d", [1, 2, 4, 8])
def test_draw_cache_changes_nothing(ctx, og, d):
    """mg_kdtree_enable_draw_cache: the cell of every stored point located once (by the same descent); leaf-level
    draws then gather a record instead of descending -- same draws, bit for bit, as before and as the oracle."""
    rng = np.random.default_rng(100 + d)
    pts = mh_like(rng, 20000, d, 0.2).clip(0.0, 1.0)
    lo, hi = np.zeros(d), np.ones(d)
    g = interpolate_pdf.InterpPdf(pts, lo, hi, ctx=ctx)
    ctx.set_seed(55)
    before = g.draw(30000)
    ctx.check(ctx.lib.mg_kdtree_enable_draw_cache(g.tree.h))
    ctx.set_seed(55)
    after = g.draw(30000)
    assert np.array_equal(before, after)
    assert np.array_equal(after, og.Tree(pts, lo, hi).draw(55, 0, 30000))
    ctx.set_seed(56)
    assert np.array_equal(g.draw_high_level(64, 2000), og.Tree(pts, lo, hi).draw(56, 0, 2000, nstop=64))   # not cached: descends
