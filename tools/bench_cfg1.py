#!/usr/bin/env python
"""BASELINE.json config 1 at its stated size: the two-model Gaussian-vs-Cauchy RJMCMC of
bin/gaussian_cauchy_efficiency.ml on its fixed 100-point data set -- per-model posterior chains
(nbin = 1e4, nskip = 100, n = 1e4, ONE chain each as the reference runs them, :99-105), Interp.make over the prior
box (:126-130), then the reversible-jump chain of 1e5 steps with interpolated jumps, pa = pb = 0.5.
The reference runs 1 chain on a CPU; here the GPU runs `--chains` independent chains of 1e5 steps and the oracle
(C++ restatement, one thread) runs ONE chain of 1e5 steps on the same trees: model fractions are compared (the GPU's
spread over chains gives the Monte Carlo error of one chain), accept rates and throughput reported.  One JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096); ap.add_argument("--steps", type=int, default=100_000)
    ap.add_argument("--seed", type=int, default=20111104)
    ap.add_argument("--posterior-chains", type=int, default=100, help="1 = the reference's single thinned chain per model")
    a = ap.parse_args()
    from mcmc_ocaml_b200 import Context, interpolate_pdf, mcmc, plugins as P
    from oracle import oracle as og
    from tests.golden.gc_data import DATA
    ctx = Context(0, a.seed)
    lo, hi = [-1.0, 0.5], [1.0, 1.5]
    prior = P.box(lo, hi, value=-0.693147)                       # bin/gaussian_cauchy_efficiency.ml:55-67
    prop = P.wrap_proposal(lo, hi, [0.1, 0.1])                   # :89-94 (Mcmc.uniform_wrapping)
    lg, lc = P.gauss_data(DATA), P.cauchy_data(DATA)
    out = {"workload": "cfg1: Gaussian-vs-Cauchy RJMCMC, 100 fixed data points, interpolated jumps from 1e4-sample posterior chains"}
    t = time.perf_counter()
    # the reference thins ONE chain (nbin = 1e4, nskip = 100, n = 1e4: 1.01e6 strictly sequential steps, 16 us each on a
    # GPU); the same 1e4 thinned samples come from 100 chains x 100 samples with the same burn-in and thinning
    pc = a.posterior_chains
    gs = mcmc.mcmc_array(10_000 // pc, lg, prior, prop, [0.0, 1.0], nchains=pc, nbin=10_000, nskip=100, ctx=ctx).values()
    cs = mcmc.mcmc_array(10_000 // pc, lc, prior, prop, [0.0, 1.0], nchains=pc, nbin=10_000, nskip=100, ctx=ctx).values()
    out["posterior_chains"], out["posterior_chains_s"] = pc, time.perf_counter() - t
    out["gaussian_posterior_mean"], out["cauchy_posterior_mean"] = gs.mean(0).tolist(), cs.mean(0).tolist()
    gi = interpolate_pdf.InterpPdf(gs, lo, hi, ctx=ctx); ci = interpolate_pdf.InterpPdf(cs, lo, hi, ctx=ctx)
    A = mcmc.RjModel(lg, prior, prop, 0.5, interp=gi); B = mcmc.RjModel(lc, prior, prop, 0.5, interp=ci)
    C, n = a.chains, a.steps
    # per-chain fractions need the per-chain model record: nskip = 100 keeps it at C x 1000 bytes
    t = time.perf_counter()
    g = mcmc.rjmcmc_array(n // 100 + 1, A, B, [0.0, 1.0], [0.0, 1.0], nskip=100, nchains=C, record_model=True, ctx=ctx)
    dt = time.perf_counter() - t
    frac_chain = (g.model == 0).mean(0)
    out.update(gpu=dict(chains=C, steps=n, seconds=dt, chain_steps_per_s=C * n / dt, gaussian_fraction=float(frac_chain.mean()),
                        gaussian_fraction_std_over_chains=float(frac_chain.std(ddof=1)), evidence_ratio=mcmc.rjmcmc_evidence_ratio(g),
                        cross_model_accept_rate=g.cross[1] / max(1, g.cross[0])))
    oa = og.rj_model(lg, prior, prop, 0.5, tree=og.Tree(gs, lo, hi)); ob = og.rj_model(lc, prior, prop, 0.5, tree=og.Tree(cs, lo, hi))
    t = time.perf_counter()
    o = og.rjmcmc_array(a.seed, 7, n, oa, ob, [0.0, 1.0], [0.0, 1.0], nchains=1, nthreads=1, record_model=False)
    odt = time.perf_counter() - t
    fo = o["counts"][0] / sum(o["counts"])
    out.update(cpu_oracle=dict(chains=1, steps=n, seconds=odt, chain_steps_per_s=n / odt, gaussian_fraction=fo, kind="port",
                               cross_model_accept_rate=o["cross"][1] / max(1, o["cross"][0]), accept_rate=o["accept"] / n,
                               note="C++ restatement of farr/mcmc-ocaml (oracle/), one thread, the reference's own size: 1 chain x 1e5 steps"))
    z = (fo - frac_chain.mean()) / frac_chain.std(ddof=1)
    out["oracle_chain_vs_gpu_ensemble_z"] = float(z)
    out["ok"] = bool(abs(z) < 4.0 and frac_chain.mean() > 0.5)
    print(json.dumps(out))


main()
