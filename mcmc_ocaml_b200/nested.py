"""Host-side mirror of the reference's ``Nested`` module (nested.mli)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .context import Context, default_context
from .plugins import LogFn


@dataclass
class NestedOutput:
    """``'a nested_output = float * float * 'a mcmc_sample array * float array`` (nested.ml:20)."""
    log_evidence: float
    log_delta_evidence: float
    points: np.ndarray          # [n][D], ascending in log-likelihood
    log_likelihood: np.ndarray  # [n]
    log_prior: np.ndarray       # [n]
    log_weights: np.ndarray     # [n]
    nlive: int
    batch: int


def nested_evidence(log_likelihood: LogFn, log_prior: LogFn, prior_low, prior_high, *, epsrel: float = 0.01,
                    nmcmc: int = 1000, nlive: int = 1000, mode_hopping_frac: float = 0.1, batch: int = 1,
                    max_points: int | None = None, ctx: Context | None = None) -> NestedOutput:
    """``Nested.nested_evidence ?epsrel ?nmcmc ?nlive ?mode_hopping_frac ...``
    (nested.ml:122-146).  ``draw_prior`` is uniform on ``[prior_low, prior_high]``
    (every caller in the reference); ``batch`` live points are replaced per
    iteration (1 = the reference's schedule)."""
    ctx = ctx or default_context()
    dim = log_likelihood.dim
    if max_points is None:
        max_points = nlive * 400
    cfg = _abi.mg_nested_cfg(dim, nlive, nmcmc, batch, epsrel, mode_hopping_frac, max_points)
    lo, hi = _abi.as_f64(prior_low), _abi.as_f64(prior_high)
    pts = np.empty((max_points, dim)); ll = np.empty(max_points); lp = np.empty(max_points); lw = np.empty(max_points)
    lev, ldev, npts = C.c_double(), C.c_double(), C.c_int64()
    ls, ps = log_likelihood.spec(), log_prior.spec()
    ctx.check(ctx.lib.mg_nested_evidence(ctx.h, C.byref(ls), C.byref(ps), _abi.ptr(lo), _abi.ptr(hi), C.byref(cfg),
                                         C.byref(lev), C.byref(ldev), C.byref(npts), _abi.ptr(pts), _abi.ptr(ll),
                                         _abi.ptr(lp), _abi.ptr(lw)))
    k = npts.value
    # views: the untouched tail of the max_points-sized buffers is never paged in
    return NestedOutput(lev.value, ldev.value, pts[:k], ll[:k], lp[:k], lw[:k], nlive, batch)


def evidence_error_and_weights(ll, nlive: int, batch: int = 1, *, ctx: Context | None = None):
    """``evidence_error_and_weights nlive all_pts`` (nested.ml:81-120) on the
    ascending log-likelihoods of all points: (log_ev, log_dev, log weights)."""
    ctx = ctx or default_context()
    ll = _abi.as_f64(ll)
    lw = np.empty(ll.size)
    lev, ldev = C.c_double(), C.c_double()
    ctx.check(ctx.lib.mg_nested_weights(ctx.h, _abi.ptr(ll), C.c_int64(ll.size), C.c_int32(nlive), C.c_int32(batch),
                                        C.byref(lev), C.byref(ldev), _abi.ptr(lw)))
    return lev.value, ldev.value, lw


def log_total_error_estimate(log_ev: float, log_dev: float, nlive: int) -> float:
    """``Nested.log_total_error_estimate`` (nested.ml:148-150)."""
    return float(_abi.load_library().mg_nested_log_total_error(log_ev, log_dev, nlive))


def posterior_samples(n: int, out: NestedOutput, *, ctx: Context | None = None) -> np.ndarray:
    """``Nested.posterior_samples n nested_output`` (nested.ml:152-178): n points drawn from the
    weighted samples by inverse CDF (``weight_binary_search_index``); repeats are expected when n
    approaches the number of points, as the reference warns."""
    ctx = ctx or default_context()
    lw = _abi.as_f64(out.log_weights)
    idx = np.empty(n, dtype=np.int64)
    ctx.check(ctx.lib.mg_nested_posterior_indices(ctx.h, _abi.ptr(lw), C.c_int64(lw.size), C.c_int64(n),
                                                  _abi.ptr(idx, _abi.c_int64_p)))
    return out.points[idx]
