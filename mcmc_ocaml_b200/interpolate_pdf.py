"""Host-side mirror of ``Interpolate_pdf.Make`` (interpolate_pdf.ml:33-63)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .context import Context
from .kd_tree import KdTree


class InterpPdf:
    """``Interpolate_pdf.make pts low high`` (interpolate_pdf.ml:111-112): a
    piecewise-constant density over the kd-tree cells of the samples."""

    def __init__(self, pts, low, high, *, ctx: Context | None = None, tree: KdTree | None = None):
        self.tree = tree if tree is not None else KdTree(pts, low, high, min_split=2, ctx=ctx)
        self.ctx = self.tree.ctx
        self.D = self.tree.D

    def _q(self, pts):
        q = _abi.as_f64(pts).reshape(-1, self.D)
        return q

    def find_cell(self, pts, nstop: int = 0) -> np.ndarray:
        """node id reached by ``find_cell`` (interpolate_pdf.ml:101-109)."""
        q = self._q(pts)
        out = np.empty(q.shape[0], np.int32)
        self.ctx.check(self.ctx.lib.mg_interp_find_cell(self.ctx.h, self.tree.h, _abi.ptr(q), C.c_int64(q.shape[0]),
                                                        C.c_int32(nstop), _abi.ptr(out, _abi.c_int32_p)))
        return out

    def jump_prob(self, pts) -> np.ndarray:
        """``jump_prob interp _ pt`` (interpolate_pdf.ml:135-142): a density, not a log."""
        return self.jump_prob_high_level(0, pts)

    def jump_prob_high_level(self, n: int, pts) -> np.ndarray:
        """``jump_prob_high_level n`` (interpolate_pdf.ml:144-159); n = 0 descends to the leaf."""
        q = self._q(pts)
        out = np.empty(q.shape[0])
        self.ctx.check(self.ctx.lib.mg_interp_jump_prob(self.ctx.h, self.tree.h, _abi.ptr(q), C.c_int64(q.shape[0]),
                                                        C.c_int32(n), _abi.ptr(out)))
        return out

    def draw(self, m: int = 1) -> np.ndarray:
        """``draw`` (interpolate_pdf.ml:114-119), m independent draws [m][D]."""
        return self.draw_high_level(0, m)

    def draw_high_level(self, n: int, m: int = 1) -> np.ndarray:
        out = np.empty((m, self.D))
        self.ctx.check(self.ctx.lib.mg_interp_draw(self.ctx.h, self.tree.h, C.c_int64(m), C.c_int32(n), _abi.ptr(out)))
        return out


def make(pts, low, high, **kw) -> InterpPdf:
    return InterpPdf(pts, low, high, **kw)
