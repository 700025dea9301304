#!/usr/bin/env python
"""GPU equivalent of the reference's bin/evidence_tool.ml: reads the output of
an MCMC on stdin (Read_write text format) and prints three estimators of the
evidence on stdout -- harmonic mean, Lebesgue, direct integration, in that
order (bin/evidence_tool.ml:42-48).  Options as in the reference: -nbox, -lebeps."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(prog="evidence_tool", prefix_chars="-")
    ap.add_argument("-nbox", type=int, default=64, help="the number of samples in each direct/Lebesgue integration box (default 64)")
    ap.add_argument("-lebeps", type=float, default=0.1, help="the truncation parameter eps in Lebesgue evidence (default 0.1)")
    ap.add_argument("-i", default="-", help="input file (default stdin)")
    a = ap.parse_args()
    from mcmc_ocaml_b200 import Context, evidence, read_write
    rows = read_write.read(a.i)
    D = rows.shape[1] - 2
    pts, ll, lp = rows[:, :D], rows[:, D], rows[:, D + 1]
    with Context(0, 0) as ctx:
        print("%g %g %g" % (evidence.evidence_harmonic_mean(ll=ll, ctx=ctx),
                            evidence.evidence_lebesgue(pts, ll, lp, n=a.nbox, eps=a.lebeps, ctx=ctx),
                            evidence.evidence_direct(pts, ll, lp, n=a.nbox, ctx=ctx)))


if __name__ == "__main__":
    main()
