"""Registered log-density and jump-proposal plugins.

The reference takes OCaml closures (mcmc.mli:58-60); closures cannot run on
the GPU, so the GPU path takes a plugin: a kind id plus a float64 parameter
blob (include/mcmc_gpu.h).  The constructors below build the blobs for the
models the reference ships in bin/ and test/.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi


class LogFn:
    """A log-likelihood or log-prior plugin (``mg_logfn``)."""

    def __init__(self, kind: int, dim: int, params=(), scale: float = 1.0):
        self.kind, self.dim, self.scale = int(kind), int(dim), float(scale)
        self.params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())

    def spec(self) -> _abi.mg_logfn:
        return _abi.mg_logfn(self.kind, self.dim, self.scale, _abi.ptr(self.params if self.params.size else None),
                             self.params.size)

    def scaled(self, s: float) -> "LogFn":
        return LogFn(self.kind, self.dim, self.params, self.scale * s)


def zero(dim):
    return LogFn(_abi.FN_ZERO, dim)


def const(dim, c):
    return LogFn(_abi.FN_CONST, dim, [c])


def box(lo, hi, value=0.0, closed=True):
    """``c`` inside ``[lo, hi]`` (closed) or ``(lo, hi)`` (open), ``-inf`` outside."""
    lo, hi = np.atleast_1d(np.asarray(lo, float)), np.atleast_1d(np.asarray(hi, float))
    return LogFn(_abi.FN_BOX_CLOSED if closed else _abi.FN_BOX_OPEN, lo.size, np.concatenate([lo, hi, [value]]))


def gauss_diag(mu, sigma):
    """``Stats.log_multi_gaussian mu sigma`` (stats.ml:103-108)."""
    mu, sigma = np.atleast_1d(np.asarray(mu, float)), np.atleast_1d(np.asarray(sigma, float))
    return LogFn(_abi.FN_GAUSS_DIAG, mu.size, np.concatenate([mu, sigma]))


def gauss_corr(mu, cov):
    """N(mu, cov) through the lower-triangular whitening matrix L with
    L^T L = cov^-1 ... stored by rows; log N = logc - |L (x - mu)|^2 / 2."""
    mu = np.atleast_1d(np.asarray(mu, float))
    cov = np.asarray(cov, float)
    d = mu.size
    chol = np.linalg.cholesky(cov)           # cov = chol chol^T
    L = np.linalg.inv(chol)                  # lower triangular, |L (x-mu)|^2 = (x-mu)^T cov^-1 (x-mu)
    packed = np.concatenate([L[i, : i + 1] for i in range(d)])
    logc = -0.5 * d * math.log(2.0 * math.pi) - float(np.sum(np.log(np.diag(chol))))
    return LogFn(_abi.FN_GAUSS_CORR, d, np.concatenate([mu, packed, [logc]]))


def gauss_data(data):
    """x = (mu, sigma); sum_i Stats.log_gaussian mu sigma data_i."""
    return LogFn(_abi.FN_GAUSS_DATA, 2, data)


def cauchy_data(data):
    """x = (x0, gamma); sum_i Stats.log_cauchy x0 gamma data_i."""
    return LogFn(_abi.FN_CAUCHY_DATA, 2, data)


def shell(centre, radius, width):
    centre = np.atleast_1d(np.asarray(centre, float))
    return LogFn(_abi.FN_SHELL, centre.size, np.concatenate([centre, [radius, width]]))


def gauss_mix(mus, sigma):
    """log sum_k exp log_multi_gaussian mu_k sigma (test/nested_test.ml:47-53)."""
    mus = np.atleast_2d(np.asarray(mus, float))
    sigma = np.atleast_1d(np.asarray(sigma, float))
    return LogFn(_abi.FN_GAUSS_MIX, mus.shape[1], np.concatenate([[mus.shape[0]], mus.ravel(), sigma]))


class Proposal:
    """A jump proposal together with its log jump probability (``mg_proposal``)."""

    def __init__(self, kind: int, dim: int, params=()):
        self.kind, self.dim = int(kind), int(dim)
        self.params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())

    def spec(self) -> _abi.mg_proposal:
        return _abi.mg_proposal(self.kind, self.dim, _abi.ptr(self.params if self.params.size else None),
                                self.params.size)


def box_proposal(h):
    """x_i + random_between (-h_i) h_i (bin/evidence_direct.ml:39-43)."""
    h = np.atleast_1d(np.asarray(h, float))
    return Proposal(_abi.PROP_BOX, h.size, h)


def wrap_proposal(lo, hi, dx):
    """``Mcmc.uniform_wrapping lo hi dx`` per coordinate (mcmc.ml:187-196)."""
    lo, hi, dx = (np.atleast_1d(np.asarray(a, float)) for a in (lo, hi, dx))
    return Proposal(_abi.PROP_WRAP, lo.size, np.concatenate([lo, hi, dx]))


def indep_gauss_proposal(mu, sigma):
    mu, sigma = np.atleast_1d(np.asarray(mu, float)), np.atleast_1d(np.asarray(sigma, float))
    return Proposal(_abi.PROP_INDEP_GAUSS, mu.size, np.concatenate([mu, sigma]))


def left_biased_proposal(sigma):
    return Proposal(_abi.PROP_LEFT_BIASED, 1, [sigma])


def one_sided_proposal(sign, width=1.0):
    """y = x + sign * width * U(0,1) (test/mcmc_test.ml:186-199)."""
    return Proposal(_abi.PROP_ONE_SIDED, 1, [float(sign), float(width)])


def combine_jump_proposals(props):
    """``Mcmc.combine_jump_proposals [(p, proposal); ...]`` (mcmc.ml:165-185): a
    mixture proposal whose log jump probability is the log-sum-exp over all components."""
    props = list(props)
    dim = props[0][1].dim
    blob = [float(len(props))]
    for w, pr in props:
        if pr.dim != dim or pr.kind == _abi.PROP_MIXTURE:
            raise _abi.InvalidArgument("combine_jump_proposals: components must be basic proposals of one dimension")
        blob += [float(w), float(pr.kind), float(pr.params.size)] + list(pr.params)
    return Proposal(_abi.PROP_MIXTURE, dim, blob)


def differential_evolution_proposal(samples, mode_hopping_frac: float = 0.0) -> Proposal:
    """``Mcmc.differential_evolution_proposal ?mode_hopping_frac to_float from_float samples`` (mcmc.ml:198-218,
    mcmc.mli:215-218) with ``to_float`` = ``from_float`` = identity on ``float array``: ``samples`` are the
    ``value``s [M][D] (or an ``McmcSamples``).  Symmetric; the scale of the ordinary mode is Gaussian with
    sigma = 2.38 / sqrt(2 D), as coded (the .mli says uniform)."""
    pts = samples.values() if hasattr(samples, "values") else _abi.as_f64(samples)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    if pts.shape[0] < 2:
        raise _abi.InvalidArgument("differential_evolution_proposal: need at least two samples")
    blob = np.concatenate([[float(mode_hopping_frac), float(pts.shape[0])], pts.ravel()])
    return Proposal(_abi.PROP_DE, pts.shape[1], blob)


def register_source(name: str, body: str, dim: int, params=(), *, ctx=None) -> LogFn:
    """Register a user log-density given as CUDA source (``mg_plugin_register_source``):
    ``body`` is the body of ``__device__ double f(const double* x, int dim, const double* p, long long np)``.
    NVRTC compiles the sampler kernel with the function inlined.  Supported by
    ``mcmc.mcmc_array`` and ``mg_logfn_eval``."""
    import ctypes as C

    from .context import default_context
    ctx = ctx or default_context()
    kind = C.c_int32()
    ctx.check(ctx.lib.mg_plugin_register_source(ctx.h, name.encode(), body.encode(), C.byref(kind)))
    return LogFn(kind.value, dim, params)
