#!/usr/bin/env python
"""BASELINE.json config 4 at batch sizes K in {1 (reference schedule, small nlive), 1024, 8192}: log Z against the
analytic value, the reference's own acceptance criterion (test/nested_test.ml:38-39: |Z - Z_true| within 2x the
reported error), and the information-based error sqrt(H / nlive) that nested sampling actually has.

The reference's error estimate (nested.ml:148-150) adds the quadrature error (high - low rectangle sums) and a flat
1 / sqrt(nlive) in quadrature; it does not contain the information H = sum w_i log L_i - log Z, so for a likelihood
that compresses the prior by H nats it underestimates the scatter of log Z by ~sqrt(H).  Running several seeds at
every K separates that from a bias of the batched shrinkage: the batched estimator is unbiased iff the mean
offset over seeds is within the seed scatter / sqrt(nseeds), and the scatter itself follows sqrt(H / nlive)."""
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
    ap.add_argument("--dim", type=int, default=16)
    ap.add_argument("--seeds", type=int, default=6)
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true", help="small sizes (test of the script itself)")
    ap.add_argument("--same-nlive", type=int, default=0, help="like for like: K = 1 against K = 64 at nlive = 1000 with this many seeds each (only these cases)")
    a = ap.parse_args()
    import numpy as np
    from scipy import integrate, special

    from mcmc_ocaml_b200 import Context, nested, plugins as P
    D, r0, w = a.dim, 2.0, 0.1
    like = P.shell(np.zeros(D), r0, w)
    prior = P.box(np.full(D, -6.0), np.full(D, 6.0), -D * math.log(12.0))
    area = 2 * math.pi ** (D / 2) / special.gamma(D / 2)
    Z, _ = integrate.quad(lambda r: area * r ** (D - 1) * math.exp(-(r - r0) ** 2 / (2 * w * w)) / math.sqrt(2 * math.pi * w * w), 0, 6)
    logZ = math.log(Z) - D * math.log(12.0)
    cases = [(1000, 1), (100_000, 1024), (100_000, 8192)]
    if a.quick:
        cases = [(200, 1), (4000, 64), (4000, 512)]
    if a.same_nlive > 0:
        cases = [(1000, 1), (1000, 64)]
    doc = {"dim": D, "log_ev_analytic": logZ, "nmcmc": 1000, "epsrel": 0.01, "cases": []}
    for nlive, K in cases:
        rows = []
        for seed in range(a.same_nlive if a.same_nlive > 0 else (min(a.seeds, 3) if K == 1 else a.seeds)):
            ctx = Context(0, 1000 + seed)
            t = time.perf_counter()
            res = nested.nested_evidence(like, prior, np.full(D, -6.0), np.full(D, 6.0), nlive=nlive, nmcmc=1000, batch=K,
                                         epsrel=0.01, max_points=nlive * 80, ctx=ctx)
            dt = time.perf_counter() - t
            wts = np.exp(res.log_weights)
            H = float(np.sum(wts * res.log_likelihood) - res.log_evidence)
            lerr = nested.log_total_error_estimate(res.log_evidence, res.log_delta_evidence, nlive)
            z, zt, err = math.exp(res.log_evidence - logZ), 1.0, math.exp(lerr - logZ)     # in units of the true Z
            rows.append({"seed": 1000 + seed, "log_ev": res.log_evidence, "offset": res.log_evidence - logZ,
                         "reported_rel_error": err, "H_nats": H, "sqrt_H_over_nlive": math.sqrt(H / nlive),
                         "within_2x_reported_error": bool(abs(z - zt) <= 2.0 * err), "seconds": dt,
                         "points": int(len(res.log_likelihood))})
            ctx.close()
        off = np.array([r["offset"] for r in rows])
        case = {"nlive": nlive, "batch": K, "runs": rows, "mean_offset": float(off.mean()),
                "scatter_of_offsets": float(off.std(ddof=1)) if len(off) > 1 else None,
                "standard_error_of_mean": float(off.std(ddof=1) / math.sqrt(len(off))) if len(off) > 1 else None,
                "expected_scatter_sqrt_H_over_nlive": float(np.mean([r["sqrt_H_over_nlive"] for r in rows])),
                "reported_rel_error_mean": float(np.mean([r["reported_rel_error"] for r in rows])),
                "fraction_within_2x_reported_error": float(np.mean([r["within_2x_reported_error"] for r in rows]))}
        doc["cases"].append(case)
        print(json.dumps({k: v for k, v in case.items() if k != "runs"}), flush=True)
    s = json.dumps(doc, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
