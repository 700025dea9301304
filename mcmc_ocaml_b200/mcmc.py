"""Host-side mirror of the reference's ``Mcmc`` module (mcmc.mli) for the GPU
path: the same entry points and argument meaning, with closures replaced by
registered plugins and one call running an ensemble of independent chains.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .context import Context, default_context
from .plugins import LogFn, Proposal


@dataclass
class McmcSamples:
    """``'a mcmc_sample array`` for every chain (mcmc.mli:33-42).

    ``block`` is [n][D+2][C] (layout "step") or [C][n][D+2] (layout "chain");
    fields 0..D-1 are ``value``, D is ``log_likelihood``, D+1 ``log_prior``.
    """
    block: np.ndarray
    dim: int
    layout: str
    accept: np.ndarray
    reject: np.ndarray

    @property
    def nchains(self):
        return self.block.shape[2] if self.layout == "step" else self.block.shape[0]

    def chain(self, c: int) -> np.ndarray:
        """[n][D+2] rows of chain c."""
        return self.block[:, :, c] if self.layout == "step" else self.block[c]

    def values(self) -> np.ndarray:
        """All samples pooled, [n*C][D]."""
        if self.layout == "step":
            return np.ascontiguousarray(self.block[:, : self.dim, :].transpose(0, 2, 1)).reshape(-1, self.dim)
        return np.ascontiguousarray(self.block[:, :, : self.dim]).reshape(-1, self.dim)

    def log_likelihood(self) -> np.ndarray:
        b = self.block[:, self.dim, :] if self.layout == "step" else self.block[:, :, self.dim]
        return np.ascontiguousarray(b).ravel()

    def log_prior(self) -> np.ndarray:
        b = self.block[:, self.dim + 1, :] if self.layout == "step" else self.block[:, :, self.dim + 1]
        return np.ascontiguousarray(b).ravel()


def _cfg(nchains, dim, nbin, nskip, n, chain_offset, x0_shared, layout):
    return _abi.mg_mcmc_cfg(nchains, dim, _abi.LAYOUT_CHAIN_MAJOR if layout == "chain" else _abi.LAYOUT_STEP_MAJOR,
                            nbin, nskip, n, chain_offset, 1 if x0_shared else 0, 0)


def mcmc_array(n: int, log_likelihood: LogFn, log_prior: LogFn, jump_proposal: Proposal, start, *,
               nbin: int = 0, nskip: int = 1, nchains: int | None = None, chain_offset: int = 0,
               layout: str = "step", ctx: Context | None = None, out: np.ndarray | None = None) -> McmcSamples:
    """``Mcmc.mcmc_array ?nbin ?nskip n log_likelihood log_prior jump_proposal
    log_jump_prob start`` (mcmc.ml:58-72) for ``nchains`` independent chains.

    ``jump_proposal`` carries its own ``log_jump_prob``.  ``start`` is one
    point [D] (shared by all chains, as in the reference) or [C][D].
    """
    ctx = ctx or default_context()
    dim = log_likelihood.dim
    x0 = _abi.as_f64(start)
    shared = x0.ndim == 1 and (nchains is None or x0.size == dim)
    if x0.ndim == 1 and not shared:
        x0 = x0.reshape(-1, dim)
    if nchains is None:
        nchains = 1 if shared else x0.shape[0]
    if not shared and x0.shape != (nchains, dim):
        raise _abi.InvalidArgument("mcmc_array: start must be [D] or [nchains][D]")
    F = dim + 2
    shape = (n, F, nchains) if layout == "step" else (nchains, n, F)
    if out is None:
        out = np.empty(shape, dtype=np.float64)
    elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise _abi.InvalidArgument("mcmc_array: bad out buffer")
    acc = np.zeros(nchains, dtype=np.int64)
    rej = np.zeros(nchains, dtype=np.int64)
    cfg = _cfg(nchains, dim, nbin, nskip, n, chain_offset, shared, layout)
    ls, ps, js = log_likelihood.spec(), log_prior.spec(), jump_proposal.spec()
    ctx.check(ctx.lib.mg_mcmc_array(ctx.h, C.byref(ls), C.byref(ps), C.byref(js), C.byref(cfg), _abi.ptr(x0),
                                    _abi.ptr(out), _abi.ptr(acc, _abi.c_int64_p), _abi.ptr(rej, _abi.c_int64_p)))
    return McmcSamples(out, dim, layout, acc, rej)


def mcmc_array_dev(log_likelihood: LogFn, log_prior: LogFn, jump_proposal: Proposal, *, nchains: int, n: int,
                   nbin: int = 0, nskip: int = 1, chain_offset: int = 0, state_ptr: int, samples_ptr: int = 0,
                   accept_ptr: int = 0, ctx: Context | None = None) -> None:
    """Device-resident form: ``state_ptr`` -> float64 [D+2][C], ``samples_ptr``
    -> float64 [n][D+2][C] (0 = record nothing), ``accept_ptr`` -> int32 [C].
    Asynchronous on the context's stream."""
    ctx = ctx or default_context()
    dim = log_likelihood.dim
    cfg = _cfg(nchains, dim, nbin, nskip, n, chain_offset, False, "step")
    ls, ps, js = log_likelihood.spec(), log_prior.spec(), jump_proposal.spec()
    ctx.check(ctx.lib.mg_mcmc_array_dev(ctx.h, C.byref(ls), C.byref(ps), C.byref(js), C.byref(cfg),
                                        C.c_void_p(state_ptr), C.c_void_p(samples_ptr), C.c_void_p(accept_ptr)))


def reset_counters(ctx: Context | None = None):
    """``Mcmc.reset_counters`` (mcmc.ml:30-32)."""
    (ctx or default_context()).reset_counters()


def get_counters(ctx: Context | None = None):
    """``Mcmc.get_counters`` (mcmc.ml:34-35)."""
    return (ctx or default_context()).get_counters()


def remove_repeat_samples(rows: np.ndarray, dim: int) -> np.ndarray:
    """``Mcmc.remove_repeat_samples (=)`` on one chain's rows [n][D+2]
    (mcmc.ml:74-81): drops a sample whose value equals its predecessor's."""
    rows = _abi.as_f64(rows)
    if rows.shape[0] == 0:
        return rows
    keep = np.ones(rows.shape[0], dtype=bool)
    keep[1:] = np.any(rows[1:, :dim] != rows[:-1, :dim], axis=1)
    return rows[keep]


# ---------------------------------------------------------------------------
# two-model reversible jump (mcmc.mli:85-182)
# ---------------------------------------------------------------------------

class RjModel:
    """One of the two models of ``make_rjmcmc_sampler`` (mcmc.ml:89-119): its
    log-likelihood, log-prior, in-model proposal, model prior ``p`` and the
    proposal used to jump INTO it (``jintoa`` / ``jintob`` with their
    ``ljpintoa`` / ``ljpintob``): either an ``InterpPdf`` (``Interp.draw`` and
    ``log (Interp.jump_prob ...)``, test/mcmc_test.ml:175-178; ``nstop`` > 0
    selects the ``*_high_level`` forms) or an independent Gaussian
    ``(mu, sigma)`` (test/mcmc_test.ml:123-127)."""

    def __init__(self, log_likelihood: LogFn, log_prior: LogFn, jump_proposal: Proposal, p: float, *,
                 interp=None, nstop: int = 0, into_gauss=None):
        self.like, self.prior, self.prop, self.p = log_likelihood, log_prior, jump_proposal, float(p)
        self.interp, self.nstop = interp, int(nstop)
        self.into_gauss = None
        if interp is None:
            if into_gauss is None:
                raise _abi.InvalidArgument("RjModel: need an InterpPdf or an independent Gaussian to jump into the model")
            mu, sigma = into_gauss
            self.into_gauss = _abi.as_f64(np.concatenate([np.atleast_1d(mu), np.atleast_1d(sigma)]))

    def spec(self) -> _abi.mg_rj_model:
        if self.interp is not None:
            into = _abi.mg_into(_abi.INTO_INTERP, self.nstop, self.interp.tree.h, _abi.c_double_p(), 0)
        else:
            into = _abi.mg_into(_abi.INTO_INDEP_GAUSS, 0, None, _abi.ptr(self.into_gauss), self.into_gauss.size)
        return _abi.mg_rj_model(self.like.spec(), self.prior.spec(), self.prop.spec(), into, self.p)


@dataclass
class RjSamples:
    model: np.ndarray | None      # uint8 [n][C], 0 = A, 1 = B
    block: np.ndarray | None      # [n][Dmax+2][C]
    counts: tuple[int, int]       # rjmcmc_model_counts over every recorded sample
    cross: tuple[int, int] = (0, 0)   # steps that proposed a jump into the other model, of which accepted


def rjmcmc_array(n: int, A: RjModel, B: RjModel, a0, b0, *, nbin: int = 0, nskip: int = 1, nchains: int = 1,
                 chain_offset: int = 0, record_model: bool = True, record_samples: bool = False,
                 ctx: Context | None = None) -> RjSamples:
    """``Mcmc.rjmcmc_array ?nbin ?nskip n (lla,llb) (lpa,lpb) (jpa,jpb) ... (pa,pb)
    (a,b)`` (mcmc.ml:121-139) for ``nchains`` independent chains."""
    ctx = ctx or default_context()
    dm = max(A.like.dim, B.like.dim)
    cfg = _abi.mg_rjmcmc_cfg(nchains, nbin, nskip, n, chain_offset, 0, 0)
    model = np.empty((n, nchains), np.uint8) if record_model else None
    block = np.empty((n, dm + 2, nchains)) if record_samples else None
    counts = (C.c_int64 * 2)()
    sa, sb = A.spec(), B.spec()
    a0, b0 = _abi.as_f64(a0), _abi.as_f64(b0)
    if a0.size != A.like.dim or b0.size != B.like.dim:      # the library reads like.dim values from each
        raise _abi.InvalidArgument("rjmcmc_array: start points must have the dimensions of their models")
    ctx.check(ctx.lib.mg_rjmcmc_array(ctx.h, C.byref(sa), C.byref(sb), C.byref(cfg), _abi.ptr(a0), _abi.ptr(b0),
                                      _abi.ptr(model, _abi.c_uint8_p), _abi.ptr(block), counts))
    cp, ca = C.c_int64(), C.c_int64()
    ctx.lib.mg_rjmcmc_jump_counters(ctx.h, C.byref(cp), C.byref(ca))
    return RjSamples(model, block, (int(counts[0]), int(counts[1])), (int(cp.value), int(ca.value)))


def rjmcmc_array_k(n: int, models, starts, *, nbin: int = 0, nskip: int = 1, nchains: int = 1, chain_offset: int = 0,
                   record_model: bool = True, record_samples: bool = False, ctx: Context | None = None) -> RjSamples:
    """k-model reversible jump (``mg_rjmcmc_array_k``): an extension of ``Mcmc.rjmcmc_array`` -- the reference's sum
    type is two-model (mcmc.ml:83-87).  ``models``: 2..8 ``RjModel``; ``starts``: one start point per model.  With two
    models the chains are those of ``rjmcmc_array``.  ``counts`` has one entry per model."""
    ctx = ctx or default_context()
    K = len(models)
    if K != len(starts):
        raise _abi.InvalidArgument("rjmcmc_array_k: one start point per model")
    dm = max(m.like.dim for m in models)
    cfg = _abi.mg_rjmcmc_cfg(nchains, nbin, nskip, n, chain_offset, 0, 0)
    model = np.empty((n, nchains), np.uint8) if record_model else None
    block = np.empty((n, dm + 2, nchains)) if record_samples else None
    counts = (C.c_int64 * K)()
    arr = (_abi.mg_rj_model * K)(*[m.spec() for m in models])
    st = [_abi.as_f64(x) for x in starts]
    for m, x in zip(models, st):
        if x.size != m.like.dim:
            raise _abi.InvalidArgument("rjmcmc_array_k: start points must have the dimensions of their models")
    sp = (_abi.c_double_p * K)(*[_abi.ptr(x) for x in st])
    ctx.check(ctx.lib.mg_rjmcmc_array_k(ctx.h, arr, C.c_int32(K), C.byref(cfg), sp, _abi.ptr(model, _abi.c_uint8_p),
                                        _abi.ptr(block), counts))
    cp, ca = C.c_int64(), C.c_int64()
    ctx.lib.mg_rjmcmc_jump_counters(ctx.h, C.byref(cp), C.byref(ca))
    return RjSamples(model, block, tuple(int(c) for c in counts), (int(cp.value), int(ca.value)))


def rjmcmc_model_counts(samples: RjSamples) -> tuple[int, int]:
    """``Mcmc.rjmcmc_model_counts`` (mcmc.ml:141-149)."""
    return samples.counts


def rjmcmc_evidence_ratio(samples: RjSamples) -> float:
    """``Mcmc.rjmcmc_evidence_ratio`` (mcmc.ml:151-153): #A / #B."""
    na, nb = samples.counts
    return float(na) / float(nb) if nb else float("inf")
