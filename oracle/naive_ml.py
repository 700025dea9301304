"""naive_ml.py -- TEST INFRASTRUCTURE (second, independent oracle).  Not part of the product.

A deliberately naive, list-based, recursive transcription of kd_tree.ml,
interpolate_pdf.ml and evidence.ml into pure Python, written directly from
the OCaml source WITHOUT looking at oracle/*.hpp|cpp: Python lists play the
OCaml lists, tuples play the records, recursion follows the `let rec`s.  Its
only purpose is to check the fast C++ restatement (liboracle.so), which in turn
checks the CUDA path -- the OCaml reference itself cannot be run here
(SURVEY.md F1).  Python floats are IEEE doubles and math.exp / math.log are
glibc's, so results are compared with liboracle.so EXACTLY
(tests/test_oracle_naive.py).  Use N <= ~2e4: every node walks its lists.

Citations are file:line in the reference tree.
"""
from __future__ import annotations

import functools
import math
import random
import sys

sys.setrecursionlimit(100000)


# ---- Pervasives.compare on floats / float arrays (no NaNs in the tests) -------------------------------------------
def compare(a, b):
    if isinstance(a, (list, tuple)):
        if len(a) != len(b):                 # structural compare of blocks: sizes first
            return -1 if len(a) < len(b) else 1
        for x, y in zip(a, b):
            c = compare(x, y)
            if c != 0:
                return c
        return 0
    return -1 if a < b else (1 if a > b else 0)   # -0.0 = 0.0


# =====================================================================================================================
# kd_tree.ml
# =====================================================================================================================
class KdTree:
    """Kd_tree.Make(O): objects are anything; coord(o) -> list of floats."""

    EMPTY = None

    def __init__(self, coord):
        self.coord = coord

    # kd_tree.ml:69-86
    def find_ith(self, comp, i, objs):
        if not objs:
            raise ValueError("find_ith: no objects")
        if len(objs) == 1:
            if i == 0:
                return objs[0]
            raise RuntimeError("find_ith: internal error")
        x, xs = objs[0], objs[1:]
        if all(comp(x, y) == 0 for y in xs):
            return x
        pvt = objs[random.randrange(len(objs))]          # List.nth objs (Random.int (List.length objs))
        lte = [o for o in objs if comp(o, pvt) <= 0]     # List.partition keeps the order
        gt = [o for o in objs if not comp(o, pvt) <= 0]
        n_lte = len(lte)
        if i < n_lte:
            return self.find_ith(comp, i, lte)
        return self.find_ith(comp, i - n_lte, gt)

    # kd_tree.ml:88-94
    def compare_along_dim(self, i, o1, o2):
        return compare(self.coord(o1)[i], self.coord(o2)[i])

    def compare_coords(self, o1, o2):
        return compare(self.coord(o1), self.coord(o2))

    # kd_tree.ml:96-110
    def bounds_of_objects(self, objs):
        if not objs:
            raise ValueError("bounds_of_objects: no objects")
        if len(objs) == 1:
            return list(self.coord(objs[0])), list(self.coord(objs[0]))
        low = list(self.coord(objs[0]))
        high = list(low)
        for x in objs[1:]:
            c = self.coord(x)
            for i in range(len(low)):
                if c[i] < low[i]:
                    low[i] = c[i]
                if c[i] > high[i]:
                    high[i] = c[i]
        return low, high

    # kd_tree.ml:112-118
    def split_bounds(self, low, high, olow, ohigh, dim):
        x = 0.5 * (self.coord(olow)[dim] + self.coord(ohigh)[dim])
        new_low, new_high = list(low), list(high)
        new_low[dim] = x
        new_high[dim] = x
        return new_low, new_high

    # kd_tree.ml:120-130
    @staticmethod
    def longest_dim(low, high):
        dim, dx_max = -1, -math.inf
        for i in range(len(low)):
            dx = high[i] - low[i]
            if dx > dx_max:
                dim, dx_max = i, dx
        return dim

    # kd_tree.ml:132-142
    @staticmethod
    def find_max(comp, objs):
        if not objs:
            raise ValueError("find_max: no objects")
        m = objs[0]
        for x in objs[1:]:
            if comp(x, m) > 0:
                m = x
        return m

    @staticmethod
    def find_min(comp, objs):
        if not objs:
            raise ValueError("find_min: no objects")
        m = objs[0]
        for x in objs[1:]:
            if comp(x, m) < 0:
                m = x
        return m

    # kd_tree.ml:144-153
    def adjust_for_empty_split(self, comp, objs, objs2):
        if not objs and not objs2:
            return [], []
        if not objs:
            mn = self.find_min(comp, objs2)
            return [o for o in objs2 if comp(o, mn) <= 0], [o for o in objs2 if not comp(o, mn) <= 0]
        if not objs2:
            mx = self.find_max(comp, objs)
            return [o for o in objs if comp(o, mx) < 0], [o for o in objs if not comp(o, mx) < 0]
        return objs, objs2

    # kd_tree.ml:155-175.  A Cell is the tuple (objs, low, high, left, right); Empty is None.
    def tree_of_objects(self, objs, low, high):
        if not objs:
            return self.EMPTY
        if len(objs) == 1:
            return (objs, low, high, self.EMPTY, self.EMPTY)
        x, xs = objs[0], objs[1:]
        if all(self.compare_coords(x, y) == 0 for y in xs):
            return (objs, low, high, self.EMPTY, self.EMPTY)
        n = len(objs)
        i = n // 2
        l, h = self.bounds_of_objects(objs)
        dim = self.longest_dim(l, h)

        def comp(a, b):
            return self.compare_along_dim(dim, a, b)

        pvt = self.find_ith(comp, i, objs)
        lte = [o for o in objs if self.compare_along_dim(dim, o, pvt) <= 0]
        gt = [o for o in objs if not self.compare_along_dim(dim, o, pvt) <= 0]
        lte, gt = self.adjust_for_empty_split(comp, lte, gt)
        lt_bound = self.find_max(comp, lte)
        gt_bound = self.find_min(comp, gt)
        new_low, new_high = self.split_bounds(low, high, lt_bound, gt_bound, dim)
        return (objs, low, high, self.tree_of_objects(lte, low, new_high), self.tree_of_objects(gt, new_low, high))

    # kd_tree.ml:177-186
    @staticmethod
    def bounds_volume(low, high):
        v = 1.0
        for i in range(len(low)):
            v = v * (high[i] - low[i])
        return v + 0.0

    def volume(self, t):
        return 0.0 if t is None else self.bounds_volume(t[1], t[2])


# =====================================================================================================================
# interpolate_pdf.ml (the deterministic parts: find_cell, jump_prob, jump_prob_high_level)
# =====================================================================================================================
class InterpolatePdf:
    def __init__(self, pts, low, high, coord=lambda p: p):
        """make pts low high (interpolate_pdf.ml:111-112); S.coord = coord"""
        self.kdt = KdTree(coord)
        self.pts = list(pts)
        self.tree = self.kdt.tree_of_objects(list(self.pts), list(low), list(high))

    # interpolate_pdf.ml:88-99
    @staticmethod
    def in_bounds(pt, low, high):
        i, n = 0, len(pt)
        while i < n and pt[i] >= low[i] and pt[i] <= high[i]:
            i += 1
        return i == n

    def in_tree(self, pt, t):
        return False if t is None else self.in_bounds(pt, t[1], t[2])

    # interpolate_pdf.ml:101-109
    def find_cell(self, pt, t=...):
        if t is ...:
            t = self.tree
        if t is None:
            raise ValueError("find_cell: empty tree")
        _, _, _, left, right = t
        if left is None and right is None:
            return t
        if self.in_tree(pt, left):
            return self.find_cell(pt, left)
        return self.find_cell(pt, right)

    # interpolate_pdf.ml:135-142
    def jump_prob(self, pt):
        n = float(len(self.pts))
        objs, low, high, _, _ = self.find_cell(pt)
        return float(len(objs)) / (self.kdt.bounds_volume(low, high) * n)

    # interpolate_pdf.ml:144-159
    def jump_prob_high_level(self, nmax, pt):
        npts = len(self.pts)

        def prob(t):
            if t is None:
                raise RuntimeError("jump_prob_high_level: encountered empty tree!")
            objs, low, high, left, right = t
            ncell = len(objs)
            if ncell <= nmax:
                return float(ncell) / (float(npts) * self.kdt.bounds_volume(low, high))
            if self.in_tree(pt, left):
                return prob(left)
            return prob(right)

        return prob(self.tree)


# =====================================================================================================================
# evidence.ml -- a sample is the tuple (coords list, ll, lp); to_coords = identity
# =====================================================================================================================
class Evidence:
    def __init__(self):
        self.kd = KdTree(lambda s: s[0])

    # evidence.ml:72-75
    @staticmethod
    def length_at_least(n, lst):
        k = 0
        while True:
            if k == len(lst):
                return n == 0
            if n <= 0:
                return True
            n -= 1
            k += 1

    # evidence.ml:83-89: List.rev_append (collect left) (collect right)
    def collect_subvolumes(self, nmax, t):
        if t is None:
            return []
        objs, _, _, left, right = t
        if not self.length_at_least(nmax, objs):
            return [t]
        return list(reversed(self.collect_subvolumes(nmax, left))) + self.collect_subvolumes(nmax, right)

    # evidence.ml:101-107
    @staticmethod
    def evidence_harmonic_mean(samples):
        linv = 0.0
        for s in samples:
            linv = linv + 1.0 / math.exp(s[1])
        return float(len(samples)) / linv

    # evidence.ml:109-120 (List.fast_sort is a stable merge sort)
    @staticmethod
    def median_sample(f, samples):
        n = len(samples)
        ssamp = sorted(samples, key=functools.cmp_to_key(lambda a, b: compare(f(a), f(b))))
        if n % 2 == 0:
            return 0.5 * (f(ssamp[n // 2 - 1]) + f(ssamp[n // 2]))
        return f(ssamp[n // 2])

    # evidence.ml:122-124
    @staticmethod
    def mean_sample(f, samples):
        n, tot = 0, 0.0
        for s in samples:
            n, tot = n + 1, tot + f(s)
        return tot / float(n)

    # evidence.ml:126-146
    @staticmethod
    def rev_remove_dups(comp, lst):
        removed = []
        rem = list(lst)
        k = 0
        while True:
            if k == len(rem):
                return removed
            if k == len(rem) - 1:
                return [rem[k]] + removed
            x, y = rem[k], rem[k + 1]
            if comp(x, y) != 0:
                removed = [x] + removed
            k += 1

    def array_to_list_remove_dups(self, samples):
        cs = lambda a, b: compare(a[0], b[0])     # compare_samples
        lsort = sorted(samples, key=functools.cmp_to_key(cs))   # List.sort: stable
        return self.rev_remove_dups(cs, lsort)

    # evidence.ml:148-165
    def evidence_direct(self, samples, n=64):
        lsamples = self.array_to_list_remove_dups(list(samples))
        low, high = self.kd.bounds_of_objects(lsamples)
        tree = self.kd.tree_of_objects(lsamples, low, high)
        integral = 0.0
        for c in self.collect_subvolumes(n, tree):
            objs = c[0]
            lo, hi = self.kd.bounds_of_objects(objs)
            vol = self.kd.bounds_volume(lo, hi)
            post = self.mean_sample(lambda s: math.exp(s[1] + s[2]), objs)
            integral = integral + vol * post
        return integral

    # evidence.ml:167-180
    @staticmethod
    def collect_samples_up_to_eps(eps, samps):
        srt = sorted(samps, key=functools.cmp_to_key(lambda a, b: compare(-a[1], -b[1])))
        collected = []
        k = 0
        while True:
            if k == len(srt):
                return collected
            if k == len(srt) - 1:
                return collected + [srt[k]]
            x, y = srt[k], srt[k + 1]
            ilx, ily = math.exp(-x[1]), math.exp(-y[1])
            delta = ily - ilx
            assert delta >= 0.0
            if delta > eps:
                return collected + [x]
            collected.append(x)
            k += 1

    # evidence.ml:182-189
    @staticmethod
    def mean_inv_like(samps):
        tot = 0.0
        for s in samps:
            tot = tot + math.exp(-s[1])
        return tot / float(len(samps))

    # evidence.ml:191-200
    @staticmethod
    def remove_dups_rev(lst):
        removed = []
        k = 0
        while True:
            if k == len(lst):
                return removed
            if k == len(lst) - 1:
                return [lst[k]] + removed
            if not (lst[k][1] == lst[k + 1][1]):
                removed = [lst[k]] + removed
            k += 1

    # evidence.ml:202-221
    def evidence_lebesgue(self, samples, n=64, eps=0.1):
        samples = self.collect_samples_up_to_eps(eps, list(samples))
        mean_il = self.mean_inv_like(samples)
        samples = self.remove_dups_rev(samples)
        low, high = self.kd.bounds_of_objects(samples)
        t = self.kd.tree_of_objects(samples, low, high)
        pm = 0.0
        ncells = 0
        for cell in self.collect_subvolumes(n, t):
            objs = cell[0]
            lo, hi = self.kd.bounds_of_objects(objs)
            vol = self.kd.bounds_volume(lo, hi)
            prior = math.exp(self.median_sample(lambda s: s[2], objs))
            pm = pm + prior * vol
            ncells += 1
        return pm / mean_il, len(samples), ncells


# ---- helpers for the comparison with liboracle.so -------------------------------------------------------------------
def flatten_bfs(tree, index_of):
    """Breadth-first flat arrays (children adjacent) of a Cell tree: split_dim, split_val, left, begin, end, perm.
    index_of(obj) -> the object's input index.  The split dim / value of a node are read off its children's boxes
    (left.high differs from the parent's high in exactly the split dimension, kd_tree.ml:174-175)."""
    if tree is None:
        return dict(split_dim=[], split_val=[], left=[], begin=[], end=[], perm=[])
    nodes = [(tree, 0)]                   # (cell, begin)
    sd, sv, left, begin, end = [], [], [], [], []
    perm = [index_of(o) for o in tree[0]]
    lvl_b, lvl_e = 0, 1
    out_perm = list(perm)
    while lvl_b < lvl_e:
        for k in range(lvl_b, lvl_e):
            cell, b = nodes[k]
            objs, low, high, l, r = cell
            begin.append(b)
            end.append(b + len(objs))
            if l is None and r is None:
                sd.append(-1); sv.append(0.0); left.append(-1)
                continue
            # the split dimension: where the left child's high was replaced (an unchanged value is possible only if
            # the split equals the inherited bound; the right child's low then tells)
            d = [i for i in range(len(low)) if l[2][i] != high[i] or r[1][i] != low[i]]
            assert len(d) == 1, d
            sd.append(d[0]); sv.append(l[2][d[0]]); left.append(len(nodes))
            nodes.append((l, b))
            nodes.append((r, b + len(l[0])))
            out_perm[b:b + len(l[0])] = [index_of(o) for o in l[0]]
            out_perm[b + len(l[0]):b + len(objs)] = [index_of(o) for o in r[0]]
        lvl_b, lvl_e = lvl_e, len(nodes)
    return dict(split_dim=sd, split_val=sv, left=left, begin=begin, end=end, perm=out_perm)


# ---------------------------------------------------------------------------------------------------------------------
# nested.ml:81-120, 148-165 and stats.ml:240-248, transcribed statement by statement (scalar Python floats are IEEE
# doubles; math.exp / log / log1p are the C library's, as OCaml's are).  Independent of oracle.cpp: compared exactly.
def log_sum_logs(a, b):
    """stats.ml:240-248"""
    import math
    if a == -math.inf and b == -math.inf:
        return -math.inf
    if b > a:
        return log_sum_logs(b, a)
    r = math.exp(b - a)
    return a + math.log1p(r)


def evidence_error_and_weights(nlive, ll):
    """nested.ml:81-120: ll = log-likelihoods of all points in ascending order.  Returns (log_ev, log_dev, wts)."""
    import math
    vol_fraction = 1.0 / float(nlive)
    log_vol_fraction = math.log(vol_fraction)
    log_reduction_frac = math.log1p(-vol_fraction)
    log_half = -0.69314718055994530942
    n = len(ll)
    wts = [-math.inf] * n
    low, high = -math.inf, -math.inf
    ilive = n - nlive
    for i in range(0, ilive):
        logli, logli1 = ll[i], ll[i + 1]
        log_dv = log_vol_fraction + (float(i) * log_reduction_frac)
        log_dlow, log_dhigh = log_dv + logli, log_dv + logli1
        low = log_sum_logs(low, log_dlow)
        high = log_sum_logs(high, log_dhigh)
        wts[i] = log_sum_logs(wts[i], log_half + log_dlow)
        wts[i + 1] = log_sum_logs(wts[i + 1], log_half + log_dhigh)
    log_dv = log_vol_fraction + (float(ilive - 1) * log_reduction_frac)
    for i in range(ilive, n):
        logli1, logli = ll[i - 1], ll[i]
        log_dlow, log_dhigh = log_dv + logli1, log_dv + logli
        low = log_sum_logs(low, log_dlow)
        high = log_sum_logs(high, log_dhigh)
        wts[i - 1] = log_sum_logs(wts[i - 1], log_half + log_dlow)
        wts[i] = log_sum_logs(wts[i], log_half + log_dhigh)
    log_ev = log_half + log_sum_logs(low, high)
    log_dev = high + math.log1p(-(math.exp(low - high)))
    wts = [w - log_ev for w in wts]
    return log_ev, log_dev, wts


def log_total_error_estimate(log_ev, log_dev, nlive):
    """nested.ml:148-150"""
    import math
    log_rel_error2 = -(math.log(float(nlive)))
    return 0.5 * log_sum_logs(2.0 * log_dev, log_rel_error2 + 2.0 * log_ev)


def weight_binary_search_index(x, running_sums):
    """nested.ml:152-165"""
    if x <= running_sums[0]:
        return 0
    ilow, ihigh = 0, len(running_sums) - 1
    while ihigh - ilow > 1:
        imid = (ilow + ihigh) // 2
        if x <= running_sums[imid]:
            ihigh = imid
        else:
            ilow = imid
    return ihigh
