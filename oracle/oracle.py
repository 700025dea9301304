"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE, see oracle.cpp).

The struct layouts are those of include/mcmc_gpu.h (shared with the product
through mcmc_ocaml_b200._abi); everything else here is independent of the
CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mcmc_ocaml_b200 import _abi
from mcmc_ocaml_b200._abi import as_f64, ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.og_last_error.restype = C.c_char_p
        for name in ("og_bounds_volume", "og_stats_mean", "og_stats_std", "og_log_sum_logs", "og_log_gaussian",
                     "og_log_cauchy", "og_log_lognormal", "og_nested_log_total_error"):
            getattr(L, name).restype = C.c_double
        L.og_remove_repeat_samples.restype = C.c_int64
        L.og_stats_mean.argtypes = [_abi.c_double_p, C.c_int64]
        L.og_stats_std.argtypes = [_abi.c_double_p, C.c_int64, C.c_int, C.c_double]
        L.og_log_sum_logs.argtypes = [C.c_double, C.c_double]
        L.og_log_gaussian.argtypes = [C.c_double] * 3
        L.og_log_cauchy.argtypes = [C.c_double] * 3
        L.og_log_lognormal.argtypes = [C.c_double] * 3
        L.og_nested_log_total_error.argtypes = [C.c_double, C.c_double, C.c_int32]
        L.og_bounds_volume.argtypes = [_abi.c_double_p, _abi.c_double_p, C.c_int32]
        L.og_kdtree_destroy.restype = None
        L.og_kdtree_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        msg = lib().og_last_error().decode()
        if rc == _abi.MG_EINVAL:
            raise _abi.InvalidArgument(msg)
        raise OracleError(msg)


U64 = C.c_uint64


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().og_philox(c, k, o)
    return list(o)


def rng_stream(seed, epoch, purpose, g, step, n):
    u = np.empty(n)
    lanes = np.empty(n, dtype=np.uint64)
    lib().og_rng_stream(U64(seed), U64(epoch), C.c_uint32(purpose), U64(g), U64(step), C.c_int(n), ptr(u),
                        lanes.ctypes.data_as(C.POINTER(C.c_uint64)))
    return u, lanes


def logfn_eval(fn, x):
    x = as_f64(x).reshape(-1, fn.dim)
    out = np.empty(x.shape[0])
    s = fn.spec()
    _check(lib().og_logfn_eval(C.byref(s), ptr(x), C.c_int64(x.shape[0]), ptr(out)))
    return out


def mcmc_array(seed, epoch, n, like, prior, prop, start, *, nbin=0, nskip=1, nchains=None, chain_offset=0,
               nthreads=1, record=True, margins=False):
    """margins=True additionally returns [n][C]: the smallest |log u - log_accept_prob| over the steps leading to
    each slot (the near-tie diagnostic of the parity tests)."""
    dim = like.dim
    x0 = as_f64(start)
    shared = x0.ndim == 1 and (nchains is None or x0.size == dim)
    if x0.ndim == 1 and not shared:
        x0 = x0.reshape(-1, dim)
    if nchains is None:
        nchains = 1 if shared else x0.shape[0]
    cfg = _abi.mg_mcmc_cfg(nchains, dim, 0, nbin, nskip, n, chain_offset, 1 if shared else 0, 0)
    out = np.empty((n, dim + 2, nchains)) if record else None
    acc = np.zeros(nchains, dtype=np.int64)
    rej = np.zeros(nchains, dtype=np.int64)
    ls, ps, js = like.spec(), prior.spec(), prop.spec()
    mg = np.full((n, nchains), np.inf) if margins else None
    _check(lib().og_mcmc_array_m(U64(seed), U64(epoch), C.byref(ls), C.byref(ps), C.byref(js), C.byref(cfg), ptr(x0),
                                 ptr(out), ptr(acc, _abi.c_int64_p), ptr(rej, _abi.c_int64_p), C.c_int(nthreads),
                                 ptr(mg)))
    if margins:
        return out, acc, rej, mg
    return out, acc, rej


class Tree:
    def __init__(self, pts, low, high, min_split=2):
        pts = as_f64(pts)
        if pts.ndim == 1:
            pts = pts.reshape(-1, 1)
        self.N, self.D = pts.shape
        self.pts = pts
        self.low, self.high = as_f64(low), as_f64(high)
        h = C.c_void_p()
        _check(lib().og_kdtree_build(ptr(pts), C.c_int64(self.N), C.c_int32(self.D), ptr(self.low), ptr(self.high),
                                     C.c_int32(min_split), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            lib().og_kdtree_destroy(self.h)
            self.h = None

    def info(self):
        n, d, nn, nl = C.c_int64(), C.c_int32(), C.c_int64(), C.c_int32()
        lib().og_kdtree_info(self.h, C.byref(n), C.byref(d), C.byref(nn), C.byref(nl))
        return dict(npoints=n.value, dim=d.value, nnodes=nn.value, nlevels=nl.value)

    def export(self):
        nn = self.info()["nnodes"]
        sd = np.empty(nn, np.int32); sv = np.empty(nn); left = np.empty(nn, np.int32)
        b = np.empty(nn, np.int32); e = np.empty(nn, np.int32); perm = np.empty(self.N, np.int32)
        lib().og_kdtree_export(self.h, ptr(sd, _abi.c_int32_p), ptr(sv), ptr(left, _abi.c_int32_p),
                               ptr(b, _abi.c_int32_p), ptr(e, _abi.c_int32_p), ptr(perm, _abi.c_int32_p))
        return dict(split_dim=sd, split_val=sv, left=left, begin=b, end=e, perm=perm)

    def find_cell(self, q, nstop=0, boxes=False):
        q = as_f64(q).reshape(-1, self.D)
        node = np.empty(q.shape[0], np.int32)
        lo = np.empty_like(q) if boxes else None
        hi = np.empty_like(q) if boxes else None
        _check(lib().og_interp_find_cell(self.h, ptr(q), C.c_int64(q.shape[0]), C.c_int32(nstop),
                                         ptr(node, _abi.c_int32_p), ptr(lo), ptr(hi)))
        return (node, lo, hi) if boxes else node

    def jump_prob(self, q, nstop=0):
        q = as_f64(q).reshape(-1, self.D)
        out = np.empty(q.shape[0])
        _check(lib().og_interp_jump_prob(self.h, ptr(q), C.c_int64(q.shape[0]), C.c_int32(nstop), ptr(out)))
        return out

    def draw(self, seed, epoch, M, nstop=0):
        out = np.empty((M, self.D))
        _check(lib().og_interp_draw(U64(seed), U64(epoch), self.h, C.c_int64(M), C.c_int32(nstop), ptr(out)))
        return out


def bounds_volume(lo, hi):
    lo, hi = as_f64(lo), as_f64(hi)
    return lib().og_bounds_volume(ptr(lo), ptr(hi), C.c_int32(lo.size))


def rj_model(like, prior, prop, p, *, tree=None, nstop=0, into_gauss=None):
    """Returns (struct, keepalive)."""
    keep = [like, prior, prop, tree, into_gauss]
    if tree is not None:
        into = _abi.mg_into(_abi.INTO_INTERP, nstop, tree.h, _abi.c_double_p(), 0)
    else:
        g = as_f64(np.concatenate([np.atleast_1d(into_gauss[0]), np.atleast_1d(into_gauss[1])]))
        keep.append(g)
        into = _abi.mg_into(_abi.INTO_INDEP_GAUSS, 0, None, ptr(g), g.size)
    return _abi.mg_rj_model(like.spec(), prior.spec(), prop.spec(), into, p), keep


def rjmcmc_array(seed, epoch, n, A, B, a0, b0, *, nbin=0, nskip=1, nchains=1, chain_offset=0, nthreads=1,
                 record_model=True, record_samples=False, margins=False):
    (ma, ka), (mb, kb) = A, B
    dm = max(ma.like.dim, mb.like.dim)
    cfg = _abi.mg_rjmcmc_cfg(nchains, nbin, nskip, n, chain_offset, 0, 0)
    model = np.empty((n, nchains), np.uint8) if record_model else None
    samples = np.empty((n, dm + 2, nchains)) if record_samples else None
    counts = (C.c_int64 * 2)()
    acc = C.c_int64()
    a0, b0 = as_f64(a0), as_f64(b0)
    mg = np.full((n, nchains), np.inf) if margins else None
    cross = (C.c_int64 * 2)()
    _check(lib().og_rjmcmc_array_m(U64(seed), U64(epoch), C.byref(ma), C.byref(mb), C.byref(cfg), ptr(a0), ptr(b0),
                                   ptr(model, _abi.c_uint8_p), ptr(samples), counts, C.byref(acc), C.c_int(nthreads),
                                   ptr(mg), cross))
    return dict(model=model, samples=samples, counts=(counts[0], counts[1]), accept=acc.value, margins=mg,
                cross=(cross[0], cross[1]))


def rjmcmc_array_k(seed, epoch, n, models, starts, *, nbin=0, nskip=1, nchains=1, chain_offset=0, nthreads=1,
                   record_model=True, record_samples=False, margins=False):
    """k-model extension (oracle.cpp: rj_chain_k).  models: list of rj_model(...) results."""
    K = len(models)
    arr = (_abi.mg_rj_model * K)(*[m for m, _ in models])
    dm = max(m.like.dim for m, _ in models)
    cfg = _abi.mg_rjmcmc_cfg(nchains, nbin, nskip, n, chain_offset, 0, 0)
    model = np.empty((n, nchains), np.uint8) if record_model else None
    samples = np.empty((n, dm + 2, nchains)) if record_samples else None
    counts = (C.c_int64 * K)()
    acc = C.c_int64()
    st = [as_f64(x) for x in starts]
    sp = (_abi.c_double_p * K)(*[ptr(x) for x in st])
    mg = np.full((n, nchains), np.inf) if margins else None
    cross = (C.c_int64 * 2)()
    _check(lib().og_rjmcmc_array_k(U64(seed), U64(epoch), arr, C.c_int32(K), C.byref(cfg), sp,
                                   ptr(model, _abi.c_uint8_p), ptr(samples), counts, C.byref(acc), C.c_int(nthreads),
                                   ptr(mg), cross))
    return dict(model=model, samples=samples, counts=tuple(counts), accept=acc.value, margins=mg,
                cross=(cross[0], cross[1]))


def evidence_harmonic_mean(ll):
    ll = as_f64(ll)
    out = (C.c_double * 2)()
    _check(lib().og_evidence_harmonic_mean(ptr(ll), C.c_int64(ll.size), out))
    return out[0], out[1]


def stats_draw(seed, epoch, kind, a, b, n):
    out = np.empty(n)
    lib().og_stats_draw(U64(seed), U64(epoch), C.c_int32(kind), C.c_double(a), C.c_double(b), C.c_int64(n), ptr(out))
    return out


def harmonic_bootstrap(seed, epoch, ll, nbstrap):
    ll = as_f64(ll)
    out = np.empty(nbstrap)
    _check(lib().og_harmonic_bootstrap(U64(seed), U64(epoch), ptr(ll), C.c_int64(ll.size), C.c_int32(nbstrap), ptr(out)))
    return out


def evidence_lebesgue(pts, ll, lp, n=64, eps=0.1, full_tree=False):
    pts, ll, lp = as_f64(pts), as_f64(ll), as_f64(lp)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    out = (C.c_double * 2)()
    nk, nc = C.c_int64(), C.c_int64()
    _check(lib().og_evidence_lebesgue(ptr(pts), ptr(ll), ptr(lp), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
                                      C.c_int32(n), C.c_double(eps), C.c_int32(1 if full_tree else 0), out,
                                      C.byref(nk), C.byref(nc)))
    return dict(value=out[0], value_ld=out[1], nkept=nk.value, ncells=nc.value)


def evidence_direct(pts, ll, lp, n=64, full_tree=False):
    pts, ll, lp = as_f64(pts), as_f64(ll), as_f64(lp)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    out = (C.c_double * 2)()
    nc = C.c_int64()
    _check(lib().og_evidence_direct(ptr(pts), ptr(ll), ptr(lp), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
                                    C.c_int32(n), C.c_int32(1 if full_tree else 0), out, C.byref(nc)))
    return dict(value=out[0], value_ld=out[1], ncells=nc.value)


def remove_repeat_samples(rows, dim):
    rows = as_f64(rows)
    out = np.empty_like(rows)
    k = lib().og_remove_repeat_samples(ptr(rows), C.c_int64(rows.shape[0]), C.c_int32(dim), ptr(out))
    return out[:k]


def mean(x):
    x = as_f64(x)
    return lib().og_stats_mean(ptr(x), x.size)


def std(x, mean=None):
    x = as_f64(x)
    return lib().og_stats_std(ptr(x), x.size, 0 if mean is None else 1, 0.0 if mean is None else mean)


def multi_mean(xs):
    xs = as_f64(xs)
    out = np.empty(xs.shape[1])
    lib().og_stats_multi_mean(ptr(xs), C.c_int64(xs.shape[0]), C.c_int32(xs.shape[1]), ptr(out))
    return out


def multi_std(xs, mean=None):
    xs = as_f64(xs)
    out = np.empty(xs.shape[1])
    m = None if mean is None else as_f64(mean)
    lib().og_stats_multi_std(ptr(xs), C.c_int64(xs.shape[0]), C.c_int32(xs.shape[1]), ptr(m), ptr(out))
    return out


def autocorrelation(x, nslides):
    x = as_f64(x)
    r = np.empty(nslides)
    L = C.c_double()
    lib().og_stats_autocorrelation(ptr(x), C.c_int64(x.size), C.c_int32(nslides), ptr(r), C.byref(L))
    return r, L.value


def log_sum_logs(a, b):
    return lib().og_log_sum_logs(a, b)


def log_gaussian(mu, sigma, x):
    return lib().og_log_gaussian(mu, sigma, x)


def log_cauchy(x0, g, x):
    return lib().og_log_cauchy(x0, g, x)


def log_lognormal(mu, sigma, x):
    return lib().og_log_lognormal(mu, sigma, x)


def draw_gaussian(seed, epoch, mu, sigma, n):
    out = np.empty(n)
    lib().og_draw_gaussian(U64(seed), U64(epoch), C.c_double(mu), C.c_double(sigma), C.c_int64(n), ptr(out))
    return out


def de_proposals(seed, epoch, table, mode_hop, z, M):
    table = as_f64(table)
    z = as_f64(z)
    out = np.empty((M, table.shape[1]))
    lib().og_de_proposals(U64(seed), U64(epoch), ptr(table), C.c_int64(table.shape[0]), C.c_int32(table.shape[1]),
                          C.c_double(mode_hop), ptr(z), C.c_int64(M), ptr(out))
    return out


def nested_evidence(seed, epoch, like, prior, prior_lo, prior_hi, *, nlive=1000, nmcmc=1000, epsrel=0.01,
                    mode_hopping_frac=0.1, batch=1, max_points=None):
    dim = like.dim
    if max_points is None:
        max_points = nlive * 200
    cfg = _abi.mg_nested_cfg(dim, nlive, nmcmc, batch, epsrel, mode_hopping_frac, max_points)
    lo, hi = as_f64(prior_lo), as_f64(prior_hi)
    pts = np.empty((max_points, dim)); ll = np.empty(max_points); lp = np.empty(max_points); lw = np.empty(max_points)
    lev, ldev, npts = C.c_double(), C.c_double(), C.c_int64()
    ls, ps = like.spec(), prior.spec()
    _check(lib().og_nested_evidence(U64(seed), U64(epoch), C.byref(ls), C.byref(ps), ptr(lo), ptr(hi), C.byref(cfg),
                                    C.byref(lev), C.byref(ldev), C.byref(npts), ptr(pts), ptr(ll), ptr(lp), ptr(lw)))
    k = npts.value
    return dict(log_ev=lev.value, log_dev=ldev.value, pts=pts[:k], ll=ll[:k], lp=lp[:k], logw=lw[:k])


def nested_weights(ll, nlive, batch=1):
    ll = as_f64(ll)
    lw = np.empty(ll.size)
    lev, ldev = C.c_double(), C.c_double()
    _check(lib().og_nested_weights(ptr(ll), C.c_int64(ll.size), C.c_int32(nlive), C.c_int32(batch), C.byref(lev),
                                   C.byref(ldev), ptr(lw)))
    return lev.value, ldev.value, lw


def nested_log_total_error(log_ev, log_dev, nlive):
    return lib().og_nested_log_total_error(log_ev, log_dev, nlive)


def nested_posterior_indices(seed, epoch, logw, n):
    logw = as_f64(logw)
    out = np.empty(n, np.int64)
    lib().og_nested_posterior_indices(U64(seed), U64(epoch), ptr(logw), C.c_int64(logw.size), C.c_int64(n),
                                      ptr(out, _abi.c_int64_p))
    return out
