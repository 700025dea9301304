"""Helpers shared by the GPU-vs-oracle parity tests."""
import numpy as np

# A GPU chain may leave the oracle's only where the accept test  log u < log_accept_prob  (mcmc.ml:47) is a near-tie:
# the two sides evaluate transcendental log-densities with different libms (CUDA vs glibc, <= 1-2 ulp apart), which can
# flip the comparison only if |log u - log_accept_prob| is of the order of that rounding difference.  Everything else
# is a bug.  1e-10 is ~1e4 times the largest libm discrepancy seen in a log-density (|ll| ~ 1e2, sums of ~1e2 terms)
# and still makes a chance coincidence (probability ~1e-10 per step) impossible at test sizes.
NEAR_TIE = 1e-10


def divergence_report(got, want, margins, what=""):
    """got / want: [n][k][C] (or [n][C]) recorded values; margins: [n][C] from the oracle (min |log u - delta| over the
    steps leading to each slot).  Returns the fraction of chains that differ anywhere, after asserting that EVERY such
    chain first differs at a slot whose accept test was a near-tie."""
    got, want = np.asarray(got), np.asarray(want)
    if got.ndim == 2:
        got, want = got[:, None, :], want[:, None, :]
    neq = np.any(got != want, axis=1)                  # [n][C]
    bad_chains = np.nonzero(neq.any(axis=0))[0]
    for c in bad_chains:
        s = int(np.argmax(neq[:, c]))                  # first differing slot
        m = float(margins[s, c])
        assert m <= NEAR_TIE, (f"{what}: chain {c} leaves the oracle at slot {s} where the accept margin is {m:.3e} "
                               f"(not a near-tie: a real divergence)")
    frac = len(bad_chains) / got.shape[2]
    print(f"[parity] {what}: {len(bad_chains)} of {got.shape[2]} chains diverge (all at near-ties); fraction {frac:.2e}")
    return frac
