"""Pin the oracle's nested.ml restatement against test/nested_test.ml."""
import math

import numpy as np
import pytest

from mcmc_ocaml_b200 import plugins as P

PRIOR = P.box([0, 0], [1, 1], 0.0, closed=False)     # nested_test.ml:24-28 (strict inequalities)


def test_single_gaussian(og):  # nested_test.ml:23-39 (with a smaller nmcmc to keep the CPU suite short)
    like = P.gauss_diag([0.5, 0.5], [0.1, 0.1])
    r = og.nested_evidence(3, 0, like, PRIOR, [0, 0], [1, 1], nlive=300, nmcmc=60)
    ev = math.exp(r["log_ev"])
    err = math.exp(og.nested_log_total_error(r["log_ev"], r["log_dev"], 300))
    assert abs(ev - 1.0) <= 2.5 * err and err < 0.15
    # weights sum to one (nested_test.ml:66-85), points ascending in ll
    assert np.exp(r["logw"]).sum() == pytest.approx(1.0, abs=1e-8)
    assert np.all(np.diff(r["ll"]) >= 0)
    mean = np.sum(np.exp(r["logw"]) * r["pts"][:, 0])
    assert mean == pytest.approx(0.5, abs=0.1)


def test_four_gaussians(og):  # nested_test.ml:41-64
    like = P.gauss_mix([[0.25, 0.25], [0.25, 0.75], [0.75, 0.25], [0.75, 0.75]], [0.05, 0.05])
    r = og.nested_evidence(4, 0, like, PRIOR, [0, 0], [1, 1], nlive=300, nmcmc=60)
    ev = math.exp(r["log_ev"])
    err = math.exp(og.nested_log_total_error(r["log_ev"], r["log_dev"], 300))
    assert abs(ev - 4.0) <= 2.5 * err and err < 0.8


def test_batched_schedule_consistent(og):
    """K-at-a-time replacement: same evidence within the error estimate; K=1
    weights reduce to the reference formula"""
    like = P.gauss_diag([0.5, 0.5], [0.1, 0.1])
    r1 = og.nested_evidence(5, 0, like, PRIOR, [0, 0], [1, 1], nlive=256, nmcmc=50, batch=1)
    r8 = og.nested_evidence(5, 0, like, PRIOR, [0, 0], [1, 1], nlive=256, nmcmc=50, batch=32)
    e1 = math.exp(og.nested_log_total_error(r1["log_ev"], r1["log_dev"], 256))
    assert abs(math.exp(r8["log_ev"]) - math.exp(r1["log_ev"])) < 4 * e1
    # reference formula for K = 1: log_dv_i = log(1/nlive) + i * log1p(-1/nlive)
    ll = r1["ll"]; n = len(ll); nlive = 256; ilive = n - nlive
    lvf, lred = math.log(1.0 / nlive), math.log1p(-1.0 / nlive)
    low = high = -math.inf
    for i in range(ilive):
        dv = lvf + i * lred
        low = og.log_sum_logs(low, dv + ll[i]); high = og.log_sum_logs(high, dv + ll[i + 1])
    dv = lvf + (ilive - 1) * lred
    for i in range(ilive, n):
        low = og.log_sum_logs(low, dv + ll[i - 1]); high = og.log_sum_logs(high, dv + ll[i])
    assert r1["log_ev"] == pytest.approx(-0.69314718055994530942 + og.log_sum_logs(low, high), rel=0, abs=1e-12)


def test_de_proposal_sqrt2(og):  # mcmc_test.ml:213-224
    rng = np.random.default_rng(0)
    table = rng.normal(10.0, 1.0, (100000, 1))
    p = og.de_proposals(1, 0, table, 1.0, [0.0], 400000)[:, 0]
    assert abs(p.mean()) < 0.01 and abs(p.std(ddof=1) - math.sqrt(2.0)) < 0.01
