#!/usr/bin/env python
"""Times BASELINE.json config 3 on one GPU: Weinberg (Lebesgue) kd-tree
evidence + harmonic mean on N synthetic D-dimensional posterior samples
(x ~ N(0.5, 0.05^2 I), ll = log_multi_gaussian, lp = 0 on the unit box,
analytic Z = 1), plus the full kd-tree build and Interpolate_pdf queries.
Inputs are device resident; times are CUDA-event / wall times after warm-up."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Args:
    def __init__(self, **kw):
        self.n, self.d, self.reps, self.dups, self.queries, self.skip_full_tree, self.device = 10_000_000, 20, 3, 0.0, 10_000_000, False, 0
        self.__dict__.update(kw)


def run(a, ctx=None):
    import ctypes as C

    import numpy as np
    import torch

    from mcmc_ocaml_b200 import Context, evidence, kd_tree
    dev = torch.device("cuda", a.device)
    ctx = ctx or Context(a.device, 12345)
    g = torch.Generator(device=dev); g.manual_seed(12345)
    N, D = a.n, a.d
    x = torch.empty((N, D), dtype=torch.float64, device=dev).normal_(0.5, 0.05, generator=g)
    if a.dups > 0:
        rep = torch.rand(N, device=dev, generator=g) < a.dups
        rep[0] = False
        idx = torch.arange(N, device=dev)
        idx[rep] = 0
        idx = torch.cummax(idx, 0).values
        x = x[idx].contiguous()
    sigma = 0.05
    ll = (-0.91893853320467274178 - math.log(sigma) - 0.5 * ((x - 0.5) / sigma) ** 2).sum(1)
    lp = torch.zeros(N, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    out = {"N": N, "D": D, "dups": a.dups}

    def timed(fn, reps=a.reps):
        fn()  # warm-up
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
            ts.append(time.perf_counter() - t)
        return min(ts), r

    t, h = timed(lambda: evidence.evidence_harmonic_mean_dev(ll.data_ptr(), N, ctx=ctx))
    out["harmonic_s"] = t; out["harmonic"] = h; out["harmonic_GBps"] = 8 * N / t / 1e9
    l0 = ctx.launch_count
    t, z = timed(lambda: evidence.evidence_lebesgue_dev(x.data_ptr(), ll.data_ptr(), lp.data_ptr(), N, D, n=64, eps=0.1, ctx=ctx))
    out["lebesgue_s"] = t; out["lebesgue_Z"] = z; out["lebesgue_samples_per_s"] = N / t
    out["lebesgue_launches"] = (ctx.launch_count - l0) // (a.reps + 1)
    t, z = timed(lambda: evidence.evidence_direct_dev(x.data_ptr(), ll.data_ptr(), lp.data_ptr(), N, D, n=64, ctx=ctx), reps=1)
    out["direct_s"] = t; out["direct_Z"] = z
    lo, hi = np.zeros(D), np.ones(D)
    def build(ms):
        tr = kd_tree.KdTree.from_device(x.data_ptr(), N, D, lo, hi, min_split=ms, ctx=ctx)
        info = (tr.nnodes, tr.nlevels)
        return tr, info
    t, (tr, info) = timed(lambda: build(64), reps=2)
    out["tree64_s"] = t; out["tree64_nodes"], out["tree64_levels"] = info
    tr.close()
    if not a.skip_full_tree:
        t, (tr, info) = timed(lambda: build(2), reps=2)
        out["tree_full_s"] = t; out["tree_full_nodes"], out["tree_full_levels"] = info
        out["tree_full_points_per_s"] = N / t
        M = a.queries
        q = torch.rand((M, D), dtype=torch.float64, device=dev, generator=g) * 0.3 + 0.35
        prob = torch.empty(M, dtype=torch.float64, device=dev)
        node = torch.empty(M, dtype=torch.int32, device=dev)
        def jp():
            ctx.check(ctx.lib.mg_interp_jump_prob_dev(ctx.h, tr.h, C.c_void_p(q.data_ptr()), C.c_int64(M), C.c_int32(0),
                                                      C.c_void_p(prob.data_ptr()), C.c_void_p(node.data_ptr())))
            ctx.sync()
        t, _ = timed(jp)
        out["jump_prob_s"] = t; out["jump_prob_queries_per_s"] = M / t
        outd = torch.empty((M, D), dtype=torch.float64, device=dev)
        def dr():
            ctx.check(ctx.lib.mg_interp_draw_dev(ctx.h, tr.h, C.c_int64(M), C.c_int32(0), C.c_void_p(outd.data_ptr())))
        t, _ = timed(dr)
        out["draw_s"] = t; out["draw_per_s"] = M / t
        out["draw_mean0"] = float(outd[:, 0].mean())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dups", type=float, default=0.0, help="fraction of rows repeating their predecessor")
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--skip-full-tree", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    print(json.dumps(run(a)))


if __name__ == "__main__":
    main()
