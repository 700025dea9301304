#!/usr/bin/env python
"""BASELINE.json config 5 on one GPU: an ensemble of two-model reversible-jump
chains between a dA-D and a dB-D isotropic Gaussian posterior, cross-model
jumps interpolated from kd-trees of ntree i.i.d. posterior draws.
Model k: flat prior on [0,1]^dk with density 1, likelihood c_k N(x; 0.5, s^2 I)
-> Z_k = c_k, so with pa = pb = 0.5 the count ratio #A/#B -> c_A / c_B."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--da", type=int, default=2)
    ap.add_argument("--db", type=int, default=4)
    ap.add_argument("--ntree", type=int, default=10_000_000)
    ap.add_argument("--ratio", type=float, default=2.0)
    ap.add_argument("--nstop", type=int, default=0, help="> 0: the *_high_level jumps (cells of <= nstop objects)")
    a = ap.parse_args()
    import numpy as np
    import torch

    from mcmc_ocaml_b200 import Context, interpolate_pdf, kd_tree, mcmc, plugins as P
    dev = torch.device("cuda", 0)
    ctx = Context(0, 20111104)
    s = 0.05
    out = {"chains": a.chains, "steps": a.steps, "dA": a.da, "dB": a.db, "ntree": a.ntree}
    g = torch.Generator(device=dev); g.manual_seed(1)
    models = []
    for d, logc in ((a.da, 0.0), (a.db, -math.log(a.ratio))):
        pts = torch.empty((a.ntree, d), dtype=torch.float64, device=dev).normal_(0.5, s, generator=g).clamp_(0.0, 1.0)
        torch.cuda.synchronize(); t = time.perf_counter()
        tree = kd_tree.KdTree.from_device(pts.data_ptr(), a.ntree, d, np.zeros(d), np.ones(d), ctx=ctx)
        torch.cuda.synchronize()
        out[f"tree_{d}d_build_s"] = time.perf_counter() - t
        out[f"tree_{d}d_nodes"] = tree.nnodes
        del pts
        interp = interpolate_pdf.InterpPdf(None, None, None, tree=tree)
        like = P.gauss_diag(np.full(d, 0.5), np.full(d, s))
        prior = P.box(np.zeros(d), np.ones(d), logc)
        prop = P.wrap_proposal(np.zeros(d), np.ones(d), np.full(d, 2.0 * s / math.sqrt(d)))
        models.append(mcmc.RjModel(like, prior, prop, 0.5, interp=interp, nstop=a.nstop))
    A, B = models
    a0, b0 = np.full(a.da, 0.5), np.full(a.db, 0.5)
    mcmc.rjmcmc_array(2, A, B, a0, b0, nskip=10, nchains=a.chains, record_model=False, ctx=ctx)   # warm-up
    torch.cuda.synchronize(); t = time.perf_counter()
    r = mcmc.rjmcmc_array(a.steps // 10 + 1, A, B, a0, b0, nskip=10, nchains=a.chains, record_model=False, ctx=ctx)
    dt = time.perf_counter() - t
    out["rj_s"] = dt
    out["rj_kernel_ms"] = ctx.last_kernel_ms
    out["chain_steps_per_s"] = a.chains * (a.steps // 10) * 10 / (ctx.last_kernel_ms * 1e-3)
    out["counts"] = r.counts
    out["ratio"] = r.counts[0] / max(1, r.counts[1])
    out["expected_ratio"] = a.ratio
    out["nstop"] = a.nstop
    out["cross_model_proposals"], out["cross_model_accepted"] = r.cross
    out["cross_model_accept_rate"] = r.cross[1] / max(1, r.cross[0])
    acc, rej = ctx.get_counters()
    out["accept_rate"] = acc / max(1, acc + rej)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
