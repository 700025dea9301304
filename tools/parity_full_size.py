#!/usr/bin/env python
"""Exact GPU-vs-oracle comparison of BASELINE.json config 3 at (up to) its full size: kd-tree of N D-dimensional
samples (every flat array bit for bit), point location / densities of a query sample, Lebesgue, direct and
harmonic-mean evidence (1e-12 relative against the oracle).  The oracle is the checker (single-threaded C++,
~2 minutes per tree at 1e7 x 20); nothing here is timed as a product number.

  python tools/parity_full_size.py --n 10000000 --d 20 --out profiles/r02_parity_cfg3_1e7.json

Data: x ~ N(0.5, 0.05^2 I) from numpy PCG64(12345) (SURVEY.md 8d cfg#3), ll = Stats.log_multi_gaussian, lp = 0;
--dups f makes a fraction f of the rows repeat their predecessor (Metropolis-Hastings-like)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

KEYS = ["split_dim", "split_val", "left", "begin", "end", "perm"]


def make_data(N, D, dups=0.0, seed=12345):
    rng = np.random.default_rng(seed)
    x = rng.normal(0.5, 0.05, (N, D))
    if dups > 0:
        rep = rng.random(N) < dups
        rep[0] = False
        idx = np.arange(N)
        idx[rep] = 0
        idx = np.maximum.accumulate(idx)
        x = np.ascontiguousarray(x[idx])
    ll = (-0.91893853320467274178 - math.log(0.05) - 0.5 * ((x - 0.5) / 0.05) ** 2).sum(1)
    lp = np.zeros(N)
    return x, ll, lp


def compare(N, D, dups=0.0, ctx=None, og=None, nquery=100000, log=print):
    """Returns a dict of findings; raises AssertionError on any mismatch."""
    from mcmc_ocaml_b200 import Context, evidence, interpolate_pdf, kd_tree
    if og is None:
        from oracle import oracle as og
    ctx = ctx or Context(0, 12345)
    out = {"N": N, "D": D, "dups": dups}
    x, ll, lp = make_data(N, D, dups)
    lo, hi = np.zeros(D), np.ones(D)
    for ms, name in ((2, "full"), (64, "min_split64")):
        t = time.perf_counter(); g = kd_tree.KdTree(x, lo, hi, min_split=ms, ctx=ctx); tg = time.perf_counter() - t
        t = time.perf_counter(); o = og.Tree(x, lo, hi, min_split=ms); to = time.perf_counter() - t
        a, b = g.export(), o.export()
        assert g.nnodes == o.info()["nnodes"] and g.nlevels == o.info()["nlevels"], name
        for k in KEYS:
            assert np.array_equal(a[k], b[k]), f"{name} tree: {k} differs"
        out[f"tree_{name}"] = {"nnodes": int(g.nnodes), "nlevels": int(g.nlevels), "identical_arrays": KEYS,
                               "gpu_call_s_incl_h2d": tg, "oracle_s": to}
        log(f"[parity] {name} tree of {N} x {D} (dups {dups}): {g.nnodes} nodes, {g.nlevels} levels, all arrays identical "
            f"(oracle {to:.1f} s)")
        if ms == 2:
            rng = np.random.default_rng(7)
            q = np.concatenate([x[rng.choice(N, nquery // 2, replace=False)],
                                rng.normal(0.5, 0.06, (nquery // 2, D))])
            ip = interpolate_pdf.InterpPdf(None, None, None, tree=g)
            assert np.array_equal(ip.find_cell(q), o.find_cell(q)), "find_cell differs"
            assert np.array_equal(ip.jump_prob(q), o.jump_prob(q)), "jump_prob differs"
            out["point_location"] = {"queries": int(len(q)), "find_cell_identical": True, "jump_prob_identical": True}
            log(f"[parity] find_cell / jump_prob of {len(q)} queries identical")
        g.close(); del o
    # evidence (the GPU builds its own truncated tree of the survivors)
    t = time.perf_counter(); ol = og.evidence_lebesgue(x, ll, lp, n=64, eps=0.1); tl = time.perf_counter() - t
    zl = evidence.evidence_lebesgue(x, ll, lp, n=64, eps=0.1, ctx=ctx)
    t = time.perf_counter(); od = og.evidence_direct(x, ll, lp, n=64); td = time.perf_counter() - t
    zd = evidence.evidence_direct(x, ll, lp, n=64, ctx=ctx)
    oh = og.evidence_harmonic_mean(ll)
    zh = evidence.evidence_harmonic_mean(ll=ll, ctx=ctx)
    rel = lambda a, b: abs(a - b) / abs(b)
    out["lebesgue"] = {"gpu": zl, "oracle_reference_order": ol["value"], "oracle_long_double": ol["value_ld"],
                       "rel_vs_reference_order": rel(zl, ol["value"]), "rel_vs_long_double": rel(zl, ol["value_ld"]),
                       "nkept": ol["nkept"], "ncells": ol["ncells"], "oracle_s": tl}
    out["direct"] = {"gpu": zd, "oracle_reference_order": od["value"], "oracle_long_double": od["value_ld"],
                     "rel_vs_reference_order": rel(zd, od["value"]), "rel_vs_long_double": rel(zd, od["value_ld"]),
                     "ncells": od["ncells"], "oracle_s": td}
    out["harmonic"] = {"gpu": zh, "oracle_reference_order": oh[0], "oracle_long_double": oh[1],
                       "rel_vs_reference_order": rel(zh, oh[0]), "rel_vs_long_double": rel(zh, oh[1])}
    log(f"[parity] lebesgue gpu {zl!r} oracle {ol['value']!r} | direct gpu {zd!r} oracle {od['value']!r} | "
        f"harmonic gpu {zh!r} oracle {oh[0]!r}")
    for k in ("lebesgue", "direct"):
        assert out[k]["rel_vs_reference_order"] <= 1e-12 and out[k]["rel_vs_long_double"] <= 1e-12, (k, out[k])
    # the harmonic mean is one 1e7-term left-to-right fold in the reference: its own rounding is ~1e-12
    assert out["harmonic"]["rel_vs_long_double"] <= 1e-12 and out["harmonic"]["rel_vs_reference_order"] <= 1e-10
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=20)
    ap.add_argument("--dups", type=float, nargs="*", default=[0.0, 0.3])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = []
    for dups in a.dups:
        res.append(compare(a.n, a.d, dups))
    doc = {"what": "GPU (libmcmcgpu.so through the C ABI) against the CPU oracle at full size: exact tree arrays, "
                   "exact point location, evidence to 1e-12", "results": res}
    s = json.dumps(doc, indent=1)
    print(s)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
