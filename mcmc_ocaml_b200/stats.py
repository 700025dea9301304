"""Host-side mirror of the ``Stats`` functions on the hot path (stats.mli:20-110)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .context import Context, default_context


def mean(xs, *, ctx: Context | None = None) -> float:
    """``Stats.mean`` (stats.ml:17-23)."""
    ctx = ctx or default_context()
    xs = _abi.as_f64(xs)
    out = C.c_double()
    ctx.check(ctx.lib.mg_stats_mean(ctx.h, _abi.ptr(xs), C.c_int64(xs.size), C.byref(out)))
    return out.value


def std(xs, mean=None, *, ctx: Context | None = None) -> float:
    """``Stats.std ?mean`` (stats.ml:35-43), n-1 in the denominator."""
    ctx = ctx or default_context()
    xs = _abi.as_f64(xs)
    out = C.c_double()
    ctx.check(ctx.lib.mg_stats_std(ctx.h, _abi.ptr(xs), C.c_int64(xs.size), C.c_int(0 if mean is None else 1),
                                   C.c_double(0.0 if mean is None else mean), C.byref(out)))
    return out.value


def multi_mean(xs, *, ctx: Context | None = None) -> np.ndarray:
    """``Stats.multi_mean`` (stats.ml:58-70); xs [n][D]."""
    ctx = ctx or default_context()
    xs = _abi.as_f64(xs)
    out = np.empty(xs.shape[1])
    ctx.check(ctx.lib.mg_stats_multi_mean(ctx.h, _abi.ptr(xs), C.c_int64(xs.shape[0]), C.c_int32(xs.shape[1]),
                                          _abi.ptr(out)))
    return out


def multi_std(xs, mean=None, *, ctx: Context | None = None) -> np.ndarray:
    """``Stats.multi_std ?mean`` (stats.ml:72-87)."""
    ctx = ctx or default_context()
    xs = _abi.as_f64(xs)
    out = np.empty(xs.shape[1])
    m = None if mean is None else _abi.as_f64(mean)
    ctx.check(ctx.lib.mg_stats_multi_std(ctx.h, _abi.ptr(xs), C.c_int64(xs.shape[0]), C.c_int32(xs.shape[1]),
                                         _abi.ptr(m), _abi.ptr(out)))
    return out


def slow_autocorrelation(nslides: int, x, *, ctx: Context | None = None) -> np.ndarray:
    """``Stats.slow_autocorrelation nslides x`` (stats.ml:223-238), lags
    0..nslides-1 (parity unpinned)."""
    return autocorrelation(x, nslides, ctx=ctx)[0]


def autocorrelation(x, nslides: int, *, ctx: Context | None = None):
    """(r[0..nslides), integrated autocorrelation length)."""
    ctx = ctx or default_context()
    x = _abi.as_f64(x)
    r = np.empty(nslides)
    L = C.c_double()
    ctx.check(ctx.lib.mg_stats_autocorrelation(ctx.h, _abi.ptr(x), C.c_int64(x.size), C.c_int32(nslides), _abi.ptr(r),
                                               C.byref(L)))
    return r, L.value


def sample_block_stats(samples_ptr: int, n: int, D: int, nchains: int, *, ctx: Context | None = None):
    """Per-field mean / std over a device sample block [n][D+2][C] (all chains pooled)."""
    ctx = ctx or default_context()
    m, s = np.empty(D + 2), np.empty(D + 2)
    ctx.check(ctx.lib.mg_stats_sample_block_dev(ctx.h, C.c_void_p(samples_ptr), C.c_int64(n), C.c_int32(D),
                                                C.c_int64(nchains), _abi.ptr(m), _abi.ptr(s)))
    return m, s


# --- draws and densities (stats.ml:89-128, 240-248) ------------------------------------------------------------

DRAW_UNIFORM, DRAW_GAUSSIAN, DRAW_CAUCHY = 0, 1, 2


def _draw(kind: int, a: float, b: float, n: int, ctx: Context | None) -> np.ndarray:
    ctx = ctx or default_context()
    out = np.empty(int(n))
    ctx.check(ctx.lib.mg_stats_draw(ctx.h, C.c_int32(kind), C.c_double(a), C.c_double(b), C.c_int64(n), _abi.ptr(out)))
    return out


def draw_uniform(a: float, b: float, n: int = 1, *, ctx: Context | None = None) -> np.ndarray:
    """``Stats.draw_uniform a b`` (stats.ml:126-128), n draws from the context's Philox stream."""
    return _draw(DRAW_UNIFORM, a, b, n, ctx)


def draw_gaussian(mu: float, sigma: float, n: int = 1, *, ctx: Context | None = None) -> np.ndarray:
    """``Stats.draw_gaussian mu sigma`` (Leva's ratio of uniforms, stats.ml:113-124)."""
    return _draw(DRAW_GAUSSIAN, mu, sigma, n, ctx)


def draw_cauchy(x0: float, gamma: float, n: int = 1, *, ctx: Context | None = None) -> np.ndarray:
    """``Stats.draw_cauchy x0 gamma`` (stats.ml:89-91)."""
    return _draw(DRAW_CAUCHY, x0, gamma, n, ctx)


def log_gaussian(mu, sigma, x):
    """``Stats.log_gaussian`` (stats.ml:98-101) on the host; the device body is the GAUSS_DIAG plugin."""
    dx = (np.asarray(x, dtype=np.float64) - mu) / sigma
    return -0.91893853320467274178 - np.log(sigma) - 0.5 * dx * dx


def log_cauchy(x0, gamma, x):
    """``Stats.log_cauchy`` (stats.ml:93-96) on the host; the device body is the CAUCHY_DATA plugin."""
    dx = (np.asarray(x, dtype=np.float64) - x0) / gamma
    return 0.0 - np.log(np.pi * gamma) - np.log(1.0 + dx * dx)


def log_multi_gaussian(mu, sigma, x) -> float:
    """``Stats.log_multi_gaussian`` (stats.ml:103-108): sequential sum over the coordinates."""
    result = 0.0
    for m, s, v in zip(np.asarray(mu, dtype=np.float64), np.asarray(sigma, dtype=np.float64), np.asarray(x, dtype=np.float64)):
        result = result + float(log_gaussian(m, s, v))
    return result + 0.0


def log_sum_logs(a: float, b: float) -> float:
    """``Stats.log_sum_logs`` (stats.ml:240-248): log(e^a + e^b) without overflow."""
    if a == -np.inf and b == -np.inf:
        return -np.inf
    if b > a:
        a, b = b, a
    return float(a + np.log1p(np.exp(b - a)))
