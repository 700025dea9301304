#!/usr/bin/env python
"""Randomised soak of the kd-tree builders against the oracle: sizes from 1 point to a few 1e5, 1 to 64 dimensions,
continuous data, wide exponent ranges, heavy exact duplicates, repeated rows, sub-ulp clusters, truncated builds.
Every flat array must be identical.  Exits non-zero on the first mismatch (the case is printed)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import numpy as np

    from mcmc_ocaml_b200 import Context, kd_tree
    from oracle import oracle as og
    ctx = Context(0, 1)
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    ncase, by_mode = 0, {}
    keys = ["split_dim", "split_val", "left", "begin", "end", "perm"]
    while time.time() < t_end:
        ncase += 1
        n = int(rng.choice([1, 2, 3, 5, 33, 257, 1000, 2047, 2048, 2049, 4097, 5000, 20000, 70000, 131072, 200001, 300000]))
        d = int(rng.choice([1, 2, 3, 5, 8, 13, 20, 33, 41, 64]))
        if n * d > 6_000_000:
            d = max(1, 6_000_000 // n)
        mode = int(rng.integers(0, 6))
        if mode == 0:
            pts = rng.normal(0.5, 0.05, (n, d))
        elif mode == 1:
            pts = rng.random((n, d)) * 10.0 ** rng.integers(-3, 4, d)
        elif mode == 2:
            pts = np.round(rng.normal(0.0, 3.0, (n, d)), int(rng.integers(0, 4)))      # heavy exact duplicates / ties
        elif mode == 3:
            pts = rng.normal(0.5, 0.05, (n, d))
            for _ in range(30):
                if n < 100:
                    break
                c0 = int(rng.integers(0, n - 80)); L = int(rng.integers(2, 80))
                pts[c0:c0 + L] = pts[c0] + rng.permutation(L)[:, None] * 2.0 ** -48
        elif mode == 4:                                                                 # MH-like repeated rows
            pts = rng.normal(0.0, 1.0, (n, d))
            rep = rng.random(n) < 0.4; rep[0] = False
            idx = np.arange(n); idx[rep] = 0; idx = np.maximum.accumulate(idx); pts = pts[idx]
        else:                                                                           # both signs, zeros, a constant column
            pts = rng.normal(0.0, 1.0, (n, d)) * 10.0 ** rng.integers(-200, 200, d)
            pts[::7, 0] = 0.0; pts[::11, 0] = -0.0
            if d > 1:
                pts[:, d - 1] = 3.25
        lo, hi = pts.min(0) - 1.0, pts.max(0) + 1.0
        ms = int(rng.choice([2, 2, 2, 8, 64, 1000]))
        t = kd_tree.KdTree(pts, lo, hi, min_split=ms, ctx=ctx)
        o = og.Tree(pts, lo, hi, min_split=ms)
        ea, eb = t.export(), o.export()
        ok = t.nnodes == o.info()["nnodes"] and t.nlevels == o.info()["nlevels"] and all(np.array_equal(ea[k], eb[k]) for k in keys)
        if not ok:
            bad = [k for k in keys if ea[k].shape != eb[k].shape or not np.array_equal(ea[k], eb[k])]
            print(json.dumps({"mismatch": dict(n=n, d=d, mode=mode, min_split=ms, case=ncase, seed=a.seed), "arrays": bad,
                              "nnodes": [int(t.nnodes), int(o.info()["nnodes"])], "nlevels": [int(t.nlevels), int(o.info()["nlevels"])]}))
            np.save("gpurun_out/stress_tree_fail_pts.npy", pts)
            sys.exit(1)
        by_mode[mode] = by_mode.get(mode, 0) + 1
        t.close()
    print(json.dumps({"ok": True, "cases": ncase, "by_mode": by_mode, "seconds": a.seconds}))


if __name__ == "__main__":
    main()
