#!/usr/bin/env python
"""Randomised parity soak: GPU (through the C ABI) against the oracle on randomly drawn shapes, for a time budget.
Targets the paths with the most machinery: the balanced MH sampler (task queue, ragged groups, nskip / nbin
combinations), the kd-tree build with the 32-bit window presort (ties, duplicates, wide exponent ranges), Lebesgue /
direct evidence, and nested sampling with the sort-and-merge replacement.  Exits non-zero on the first mismatch."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import numpy as np

    from mcmc_ocaml_b200 import Context, evidence, kd_tree, mcmc, nested, plugins as P
    from oracle import oracle as og
    ctx = Context(0, 1)
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    counts = {"mh": 0, "tree": 0, "evidence": 0, "nested": 0}
    spent = {"mh": 0.0, "tree": 0.0, "evidence": 0.0, "nested": 0.0}
    it = 0
    while time.time() < t_end:
        it += 1
        kind = ["mh", "tree", "evidence", "nested"][it % 4]
        t_case = time.time()
        if kind == "mh":
            D = int(rng.choice([2, 3, 4, 7, 10, 12, 16]))
            C = int(rng.integers(19000, 26000))
            nskip = int(rng.choice([1, 1, 2, 5, 130]))
            nbin = int(rng.choice([0, 1, 127, 128, 129, 300]))
            n = int(rng.integers(2, 5)) if nskip == 130 else int(rng.integers(200, 420) // max(1, nskip // 2 + 1)) + 1
            if nbin + (n - 1) * nskip < 512:
                n = (512 - nbin) // nskip + 2
            mu = np.arange(D) / 10.0
            cov = 0.7 ** np.abs(np.subtract.outer(np.arange(D), np.arange(D)))
            like, prior, prop = P.gauss_corr(mu, cov), P.zero(D), P.box_proposal(np.full(D, 0.5))
            seed = int(rng.integers(1, 2 ** 40))
            ctx.set_seed(seed)
            got = mcmc.mcmc_array(n, like, prior, prop, mu, nchains=C, nbin=nbin, nskip=nskip, ctx=ctx)
            want, acc, rej = og.mcmc_array(seed, 0, n, like, prior, prop, mu, nchains=C, nbin=nbin, nskip=nskip, nthreads=16)
            ok = np.array_equal(got.block, want) and np.array_equal(got.accept, acc)
            desc = dict(D=D, C=C, nbin=nbin, nskip=nskip, n=n, seed=seed)
        elif kind == "tree":
            n = int(rng.choice([70000, 100003, 131072, 200000]))
            d = int(rng.choice([1, 2, 3, 5, 8]))
            mode = int(rng.integers(0, 4))
            if mode == 0:
                pts = rng.normal(0.5, 0.05, (n, d))
            elif mode == 1:
                pts = rng.random((n, d)) * 10.0 ** rng.integers(-3, 4, d)
            elif mode == 2:
                pts = np.round(rng.normal(0.0, 3.0, (n, d)), int(rng.integers(1, 4)))      # heavy exact duplicates
            else:
                pts = rng.normal(0.5, 0.05, (n, d))
                for _ in range(30):
                    c0 = int(rng.integers(0, n - 80)); L = int(rng.integers(2, 80))
                    pts[c0:c0 + L] = pts[c0] + rng.permutation(L)[:, None] * 2.0 ** -48
            lo, hi = pts.min(0) - 1.0, pts.max(0) + 1.0
            ms = int(rng.choice([2, 2, 64]))
            t = kd_tree.KdTree(pts, lo, hi, min_split=ms, ctx=ctx)
            o = og.Tree(pts, lo, hi, min_split=ms)
            ea, eb = t.export(), o.export()
            ok = t.nnodes == o.info()["nnodes"] and all(np.array_equal(ea[k], eb[k]) for k in ["split_dim", "split_val", "left", "begin", "end", "perm"])
            desc = dict(n=n, d=d, mode=mode, min_split=ms)
        elif kind == "evidence":
            n = int(rng.choice([70000, 90001, 150000]))
            d = int(rng.choice([2, 3, 6]))
            x = rng.normal(0.5, 0.05, (n, d))
            if rng.random() < 0.5:
                rep = rng.random(n) < 0.3; rep[0] = False
                idx = np.arange(n); idx[rep] = 0; idx = np.maximum.accumulate(idx); x = x[idx]
            ll = (-0.91893853320467274178 - np.log(0.05) - 0.5 * ((x - 0.5) / 0.05) ** 2).sum(1)
            lp = np.zeros(n)
            g1 = evidence.evidence_lebesgue(x, ll, lp, n=64, eps=0.1, ctx=ctx)
            o1 = og.evidence_lebesgue(x, ll, lp, 64, 0.1)["value"]
            ok = abs(g1 - o1) <= 1e-12 * abs(o1)
            desc = dict(n=n, d=d, gpu=g1, oracle=o1)
        else:
            D = int(rng.choice([2, 3, 5, 8]))
            nlive = int(rng.integers(100, 400)); batch = int(rng.integers(1, nlive // 3)); nmcmc = int(rng.integers(5, 30))
            like = P.shell(np.full(D, 0.5), 0.3, 0.05) if rng.random() < 0.5 else P.gauss_diag(np.full(D, 0.5), np.full(D, 0.1))
            prior = P.box(np.zeros(D), np.ones(D), 0.0)
            seed = int(rng.integers(1, 2 ** 40))
            ctx.set_seed(seed)
            g = nested.nested_evidence(like, prior, np.zeros(D), np.ones(D), nlive=nlive, nmcmc=nmcmc, batch=batch, ctx=ctx)
            o = og.nested_evidence(seed, 0, like, prior, np.zeros(D), np.ones(D), nlive=nlive, nmcmc=nmcmc, batch=batch)
            ok = len(g.log_likelihood) == len(o["ll"]) and np.array_equal(g.points, o["pts"])
            desc = dict(D=D, nlive=nlive, batch=batch, nmcmc=nmcmc, seed=seed)
        if not ok:
            print(json.dumps({"mismatch": kind, "case": desc, "iteration": it}))
            sys.exit(1)
        counts[kind] += 1
        spent[kind] += time.time() - t_case
    print(json.dumps({"ok": True, "seconds": a.seconds, "cases": counts, "seconds_per_kind": {k: round(v, 1) for k, v in spent.items()}}))


if __name__ == "__main__":
    main()
