"""CPU restatement of farr/mcmc-ocaml's ellipse.ml -- TEST INFRASTRUCTURE, not part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.  Every function cites the
reference lines it follows.  Arithmetic order is the reference's: the left folds of `center` and `sigma2` (one
division per term) are reproduced with numpy's sequential `cumsum`, `elliptical_range` with explicit loops over the
dimensions.  The eigen-decomposition is `numpy.linalg.eigh` (LAPACK dsyevd); the reference calls LAPACK dsyevr through
Lacaml (ellipse.ml:58-61), which is absent here: same eigenvalues in ascending order up to rounding, eigenvectors up
to sign.  **Parity of the eigen-system is therefore pinned to 1e-12, not bit for bit**; the reference's own test
(ellipse_test.ml:60-71) asks for the reconstruction of the matrix only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Ellipse:                      # ellipse.ml:21-24
    center: np.ndarray
    axes: np.ndarray
    orientation: np.ndarray         # orientation[i][j] = component i of eigenvector j


@dataclass
class Tree:                         # ellipse.ml:26-34 ('a ellipse_tree_data); Empty = None
    pts: np.ndarray                 # ids of the node's points in input order
    left: "Tree | None"
    right: "Tree | None"
    ellipse: Ellipse
    circumcircle: tuple = field(default=None)


def center(pts: np.ndarray) -> np.ndarray:
    """ellipse.ml:36-46: cen.(j) <- cen.(j) +. pt.(j) /. nf, left to right"""
    nf = float(pts.shape[0])
    return np.cumsum(pts / nf, axis=0)[-1].copy()


def sigma2(mu: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """ellipse.ml:48-66: sigma.(j).(k) <- sigma.(j).(k) +. dxj *. dxk /. nf, left to right"""
    n, d = pts.shape
    nf = float(n)
    dx = pts - mu
    s = np.zeros((d, d))
    for j in range(d):
        for k in range(j, d):
            v = np.cumsum(dx[:, j] * dx[:, k] / nf)[-1]
            s[j, k] = v
            s[k, j] = v
    return s


def eigensystem(sigma: np.ndarray):
    """ellipse.ml:58-61 (Lacaml syevr ~vectors:true): ascending eigenvalues, eigenvectors in the columns"""
    w, z = np.linalg.eigh(sigma)
    return w, z


def elliptical_range(ell: Ellipse, pt: np.ndarray) -> float:
    """ellipse.ml:63-73"""
    c, a, ori = ell.center, ell.axes, ell.orientation
    r = 0.0
    for j in range(ori.shape[1]):
        d = 0.0
        for i in range(ori.shape[0]):
            d = d + (pt[i] - c[i]) * ori[i][j]
        r = r + d * d / a[j]
    return r + 0.0


def elliptical_ranges(ell: Ellipse, pts: np.ndarray) -> np.ndarray:
    """elliptical_range of every row, same operation order, vectorised over the points only"""
    c, a, ori = ell.center, ell.axes, ell.orientation
    r = np.zeros(pts.shape[0])
    for j in range(ori.shape[1]):
        d = np.zeros(pts.shape[0])
        for i in range(ori.shape[0]):
            d = d + (pts[:, i] - c[i]) * ori[i][j]
        r = r + d * d / a[j]
    return r + 0.0


def max_elliptical_range(ell: Ellipse, pts: np.ndarray) -> float:
    """ellipse.ml:75-81: fold max from neg_infinity (Pervasives.max a b = if a >= b then a else b)"""
    m = -math.inf
    for r in elliptical_ranges(ell, pts):
        m = m if m >= r else r
    return m


def rescale_ellipse(sf: float, ell: Ellipse, pts: np.ndarray) -> Ellipse:
    """ellipse.ml:83-86"""
    r_max = max_elliptical_range(ell, pts)
    dim_sf = sf ** (1.0 / float(len(ell.axes)))
    return Ellipse(ell.center, np.array([x * dim_sf * r_max for x in ell.axes]), ell.orientation)


def enclosing_ellipse(sf: float, pts: np.ndarray) -> Ellipse:
    """ellipse.ml:98-103"""
    cen = center(pts)
    evals, evecs = eigensystem(sigma2(cen, pts))
    return rescale_ellipse(sf, Ellipse(cen, evals, evecs), pts)


def distance(p1, p2) -> float:
    """ellipse.ml:105-110"""
    r = 0.0
    for i in range(len(p1)):
        dx = p1[i] - p2[i]
        r = r + dx * dx
    return math.sqrt(r)


def union_circumcircles(cc1, cc2):
    """ellipse.ml:112-127"""
    (c1, r1), (c2, r2) = cc1, cc2
    r12 = distance(c1, c2)
    if r12 + r2 < r1:
        return cc1
    if r12 + r1 < r2:
        return cc2
    rnew = r1 + r12 + r2
    mag = 0.5 * (r2 + r12 - r1) / r12
    return (np.array([c1[i] + mag * (c2[i] - c1[i]) for i in range(len(c1))]), rnew)


def in_circumcircle(pt, cc) -> bool:
    """ellipse.ml:129-131"""
    return distance(pt, cc[0]) < cc[1]


def ellipse_circumcircle(ell: Ellipse):
    """ellipse.ml:133-134: (center, fold max neg_infinity axes)"""
    m = -math.inf
    for a in ell.axes:
        m = max(m, a)
    return (ell.center, m)


def widest_dimension(ell: Ellipse) -> int:
    """ellipse.ml:136-145: first index of the strictly largest axis"""
    imax, rmax = -1, -math.inf
    for i, a in enumerate(ell.axes):
        if a > rmax:
            rmax, imax = a, i
    return imax


def ellipse_tree(sf: float, pts: np.ndarray, ids: np.ndarray | None = None, depth: int = 0) -> Tree:
    """ellipse.ml:150-173.  `ids` are the input rows of this node (the reference carries the points themselves)."""
    if ids is None:
        ids = np.arange(pts.shape[0])
    sub = pts[ids]
    ell = enclosing_ellipse(sf, sub)
    ndim = len(ell.axes)
    assert len(ids) >= ndim + 1                                   # :152
    if depth > 900:
        raise RecursionError("ellipse_tree: unsplittable node (the reference overflows its stack)")
    split = widest_dimension(ell)
    mask = sub[:, split] < ell.center[split]                      # :154-156, List.partition keeps the order
    left_ids, right_ids = ids[mask], ids[~mask]
    if len(left_ids) == len(ids) or len(right_ids) == len(ids):
        raise RecursionError("ellipse_tree: unsplittable node (the reference overflows its stack)")
    left = None if len(left_ids) < ndim + 1 else ellipse_tree(sf, pts, left_ids, depth + 1)      # :157
    right = None if len(right_ids) < ndim + 1 else ellipse_tree(sf, pts, right_ids, depth + 1)   # :158
    own = ellipse_circumcircle(ell)
    if left is None and right is None:                            # :159-167
        cc = own
    elif left is not None and right is not None:
        cc = union_circumcircles(union_circumcircles(left.circumcircle, right.circumcircle), own)
    else:
        cc = union_circumcircles((left or right).circumcircle, own)
    return Tree(ids, left, right, ell, cc)


def flatten(tree: Tree):
    """Breadth-first node list (children of a level in parent order, left before right): the numbering of the GPU
    tree.  Returns dict of arrays: left, right (-1 = Empty), ids (list of arrays), center, axes, orientation,
    cc_center, cc_radius."""
    nodes, level = [], [tree]
    while level:
        nodes.extend(level)
        level = [c for t in level for c in (t.left, t.right) if c is not None]
    index = {id(t): i for i, t in enumerate(nodes)}
    return dict(
        left=np.array([index[id(t.left)] if t.left is not None else -1 for t in nodes], np.int32),
        right=np.array([index[id(t.right)] if t.right is not None else -1 for t in nodes], np.int32),
        ids=[t.pts for t in nodes],
        center=np.array([t.ellipse.center for t in nodes]),
        axes=np.array([t.ellipse.axes for t in nodes]),
        orientation=np.array([t.ellipse.orientation for t in nodes]),
        cc_center=np.array([t.circumcircle[0] for t in nodes]),
        cc_radius=np.array([t.circumcircle[1] for t in nodes]),
    )
