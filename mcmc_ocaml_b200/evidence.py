"""Host-side mirror of ``Evidence.Make`` (evidence.ml:31-58) for the GPU path.
Samples are given as coordinate rows [N][D] with their ``log_likelihood`` and
``log_prior`` (the fields of ``'a Mcmc.mcmc_sample``), or as ``McmcSamples``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .context import Context, default_context
from .kd_tree import KdTree


def _unpack(samples, ll=None, lp=None):
    if hasattr(samples, "values") and ll is None:          # McmcSamples
        return _abi.as_f64(samples.values()), _abi.as_f64(samples.log_likelihood()), _abi.as_f64(samples.log_prior())
    pts = _abi.as_f64(samples)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    ll, lp = _abi.as_f64(ll), _abi.as_f64(lp)
    if ll.ndim and lp.ndim and (ll.size != pts.shape[0] or lp.size != pts.shape[0]):   # the library reads N of each
        raise _abi.InvalidArgument("evidence: log_likelihood / log_prior must have one entry per sample")
    return pts, ll, lp


def kd_tree_of_samples(samples, low, high, *, ctx: Context | None = None) -> KdTree:
    """``Evidence.kd_tree_of_samples`` (evidence.ml:70)."""
    pts, _, _ = _unpack(samples, 0, 0) if not hasattr(samples, "values") else _unpack(samples)
    return KdTree(pts, low, high, ctx=ctx)


def evidence_harmonic_mean(samples=None, ll=None, *, ctx: Context | None = None) -> float:
    """``Evidence.evidence_harmonic_mean`` (evidence.ml:101-107)."""
    ctx = ctx or default_context()
    if ll is None:
        ll = samples.log_likelihood()
    ll = _abi.as_f64(ll)
    out = C.c_double()
    ctx.check(ctx.lib.mg_evidence_harmonic_mean(ctx.h, _abi.ptr(ll), C.c_int64(ll.size), C.byref(out)))
    return out.value


def harmonic_bootstrap(ll, nbstrap: int = 10000, *, ctx: Context | None = None) -> np.ndarray:
    """Bootstrap replicates of the harmonic-mean evidence (bin/harmonic_evidence.ml:41-52),
    sorted ascending like the reference's ``Array.fast_sort``."""
    ctx = ctx or default_context()
    ll = _abi.as_f64(ll)
    out = np.empty(nbstrap)
    ctx.check(ctx.lib.mg_evidence_harmonic_bootstrap(ctx.h, _abi.ptr(ll), C.c_int64(ll.size), C.c_int32(nbstrap), _abi.ptr(out)))
    return np.sort(out)


def evidence_lebesgue(samples, ll=None, lp=None, *, n: int = 64, eps: float = 0.1, ctx: Context | None = None) -> float:
    """``Evidence.evidence_lebesgue ?n ?eps`` (evidence.ml:202-221), Weinberg's
    Lebesgue integral of 1/L over kd-tree cells."""
    ctx = ctx or default_context()
    pts, ll, lp = _unpack(samples, ll, lp)
    out = C.c_double()
    ctx.check(ctx.lib.mg_evidence_lebesgue(ctx.h, _abi.ptr(pts), _abi.ptr(ll), _abi.ptr(lp), C.c_int64(pts.shape[0]),
                                           C.c_int32(pts.shape[1]), C.c_int32(n), C.c_double(eps), C.byref(out)))
    return out.value


def evidence_direct(samples, ll=None, lp=None, *, n: int = 64, ctx: Context | None = None) -> float:
    """``Evidence.evidence_direct ?n`` (evidence.ml:148-165)."""
    ctx = ctx or default_context()
    pts, ll, lp = _unpack(samples, ll, lp)
    out = C.c_double()
    ctx.check(ctx.lib.mg_evidence_direct(ctx.h, _abi.ptr(pts), _abi.ptr(ll), _abi.ptr(lp), C.c_int64(pts.shape[0]),
                                         C.c_int32(pts.shape[1]), C.c_int32(n), C.byref(out)))
    return out.value


def evidence_lebesgue_dev(pts_ptr: int, ll_ptr: int, lp_ptr: int, N: int, D: int, *, n: int = 64, eps: float = 0.1,
                          ctx: Context | None = None) -> float:
    """Device-resident form: pointers to float64 [N][D], [N], [N] on the context's GPU."""
    ctx = ctx or default_context()
    out = C.c_double()
    ctx.check(ctx.lib.mg_evidence_lebesgue_dev(ctx.h, C.c_void_p(pts_ptr), C.c_void_p(ll_ptr), C.c_void_p(lp_ptr),
                                               C.c_int64(N), C.c_int32(D), C.c_int32(n), C.c_double(eps), C.byref(out)))
    return out.value


def evidence_direct_dev(pts_ptr: int, ll_ptr: int, lp_ptr: int, N: int, D: int, *, n: int = 64,
                        ctx: Context | None = None) -> float:
    ctx = ctx or default_context()
    out = C.c_double()
    ctx.check(ctx.lib.mg_evidence_direct_dev(ctx.h, C.c_void_p(pts_ptr), C.c_void_p(ll_ptr), C.c_void_p(lp_ptr),
                                             C.c_int64(N), C.c_int32(D), C.c_int32(n), C.byref(out)))
    return out.value


def evidence_harmonic_mean_dev(ll_ptr: int, N: int, *, ctx: Context | None = None) -> float:
    ctx = ctx or default_context()
    out = C.c_double()
    ctx.check(ctx.lib.mg_evidence_harmonic_mean_dev(ctx.h, C.c_void_p(ll_ptr), C.c_int64(N), C.byref(out)))
    return out.value
