#!/usr/bin/env python
"""GPU equivalent of the reference's bin/harmonic_evidence.ml: harmonic-mean
evidence of a chain with bootstrap error bars (bin/harmonic_evidence.ml:35-53).
Options as in the reference: -nbstrap, -i, -seed."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(prog="harmonic_evidence")
    ap.add_argument("-nbstrap", type=int, default=10000, help="number of bootstrap samples to use to estimate error (default 10000)")
    ap.add_argument("-i", default="-", help="input filename")
    ap.add_argument("-seed", type=int, default=0, help="seed the RNG used for bootstrap")
    a = ap.parse_args()
    from mcmc_ocaml_b200 import Context, evidence, read_write
    rows = read_write.read(a.i)
    ll = rows[:, -2]
    with Context(0, a.seed) as ctx:
        ev = evidence.evidence_harmonic_mean(ll=ll, ctx=ctx)
        evs = evidence.harmonic_bootstrap(ll, a.nbstrap, ctx=ctx)
    ilow, ihigh = a.nbstrap // 20, (a.nbstrap * 19) // 20
    print("    Best        10%%         90%%    \n%10g %10g %10g" % (ev, evs[ilow], evs[ihigh]))


if __name__ == "__main__":
    main()
