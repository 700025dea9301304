"""Host-side mirror of the reference's ``Ellipse`` module (ellipse.ml) over the C ABI (``mg_ellipse_*``).

Names and argument order follow the reference with ``to_coord`` fixed to the identity on float rows (what every
caller uses): ``enclosing_ellipse sf pts``, ``elliptical_range ell pt``, ``ellipse_tree sf pts``,
``in_circumcircle pt cc``.  The tree is returned as flat breadth-first arrays (``EllipseTree``)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .context import Context, default_context

c_int32_p = C.POINTER(C.c_int32)


@dataclass
class Ellipse:
    """``type ellipse = {center; axes; orientation}`` (ellipse.ml:21-24); ``orientation[i][j]`` is component i of
    eigenvector j, ``axes`` the scaled eigenvalues in ascending order."""
    center: np.ndarray
    axes: np.ndarray
    orientation: np.ndarray


def enclosing_ellipse(sf: float, pts, *, ctx: Context | None = None) -> Ellipse:
    """``Ellipse.enclosing_ellipse sf to_coord pts`` (ellipse.ml:98-103)."""
    ctx = ctx or default_context()
    pts = _abi.as_f64(pts)
    if pts.ndim != 2 or pts.shape[0] < 1:
        raise _abi.InvalidArgument("enclosing_ellipse: pts must be [N][D] with N >= 1")
    n, d = pts.shape
    cen, axes, ori = np.empty(d), np.empty(d), np.empty((d, d))
    ctx.check(ctx.lib.mg_ellipse_enclosing(ctx.h, _abi.ptr(pts), C.c_int64(n), C.c_int32(d), C.c_double(sf),
                                           _abi.ptr(cen), _abi.ptr(axes), _abi.ptr(ori)))
    return Ellipse(cen, axes, ori)


def elliptical_range(ell: Ellipse, pt, *, ctx: Context | None = None):
    """``Ellipse.elliptical_range ell pt`` (ellipse.ml:63-73) for one point ([D]) or many ([M][D])."""
    ctx = ctx or default_context()
    q = _abi.as_f64(pt)
    one = q.ndim == 1
    q = np.ascontiguousarray(q.reshape(-1, ell.center.size))
    out = np.empty(q.shape[0])
    ctx.check(ctx.lib.mg_ellipse_range(ctx.h, _abi.ptr(_abi.as_f64(ell.center)), _abi.ptr(_abi.as_f64(ell.axes)),
                                       _abi.ptr(_abi.as_f64(ell.orientation)), C.c_int32(ell.center.size), _abi.ptr(q),
                                       C.c_int64(q.shape[0]), _abi.ptr(out)))
    return float(out[0]) if one else out


def in_circumcircle(pt, cc) -> bool:
    """``Ellipse.in_circumcircle pt (c, r)`` (ellipse.ml:129-131): host arithmetic, one point."""
    c, r = cc
    d = 0.0
    for a, b in zip(np.asarray(pt, float), np.asarray(c, float)):
        d = d + (a - b) * (a - b)
    return bool(np.sqrt(d) < r)


class EllipseTree:
    """``Ellipse.ellipse_tree sf to_coord pts`` (ellipse.ml:150-173) as flat arrays, nodes numbered breadth first:
    ``left`` / ``right`` (-1 = ``Empty``), ``begin`` / ``end`` into ``perm`` (the node's points as a set;
    ``points_of(k)`` lists them in input order like the reference's ``pts`` field), the node's ``ellipse``
    (``center``, ``axes``, ``orientation``) and ``circumcircle`` (``cc_center``, ``cc_radius``)."""

    def __init__(self, sf: float, pts=None, *, device_ptr: int | None = None, n: int | None = None, dim: int | None = None,
                 ctx: Context | None = None):
        self.ctx = ctx or default_context()
        lib = self.ctx.lib
        h = C.c_void_p()
        if device_ptr is not None:
            self.ctx.check(lib.mg_ellipse_tree_build_dev(self.ctx.h, C.c_void_p(device_ptr), C.c_int64(n), C.c_int32(dim),
                                                         C.c_double(sf), C.byref(h)))
        else:
            pts = _abi.as_f64(pts)
            if pts.ndim != 2:
                raise _abi.InvalidArgument("ellipse_tree: pts must be [N][D]")
            self.ctx.check(lib.mg_ellipse_tree_build(self.ctx.h, _abi.ptr(pts), C.c_int64(pts.shape[0]),
                                                     C.c_int32(pts.shape[1]), C.c_double(sf), C.byref(h)))
        self.h = h
        npts, d, nn, nl = C.c_int64(), C.c_int32(), C.c_int64(), C.c_int32()
        self.ctx.check(lib.mg_ellipse_tree_info(h, C.byref(npts), C.byref(d), C.byref(nn), C.byref(nl)))
        self.npoints, self.dim, self.nnodes, self.nlevels = npts.value, d.value, nn.value, nl.value
        self._arrays = None

    def export(self) -> dict:
        if self._arrays is None:
            n, d = self.nnodes, self.dim
            a = dict(left=np.empty(n, np.int32), right=np.empty(n, np.int32), begin=np.empty(n, np.int32),
                     end=np.empty(n, np.int32), perm=np.empty(self.npoints, np.int32), center=np.empty((n, d)),
                     axes=np.empty((n, d)), orientation=np.empty((n, d, d)), cc_center=np.empty((n, d)),
                     cc_radius=np.empty(n))
            ip = lambda x: x.ctypes.data_as(c_int32_p)
            self.ctx.check(self.ctx.lib.mg_ellipse_tree_export(
                self.h, ip(a["left"]), ip(a["right"]), ip(a["begin"]), ip(a["end"]), ip(a["perm"]), _abi.ptr(a["center"]),
                _abi.ptr(a["axes"]), _abi.ptr(a["orientation"]), _abi.ptr(a["cc_center"]), _abi.ptr(a["cc_radius"])))
            self._arrays = a
        return self._arrays

    def points_of(self, node: int) -> np.ndarray:
        a = self.export()
        return np.sort(a["perm"][a["begin"][node]:a["end"][node]])

    def ellipse_of(self, node: int) -> Ellipse:
        a = self.export()
        return Ellipse(a["center"][node], a["axes"][node], a["orientation"][node])

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.mg_ellipse_tree_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ellipse_tree(sf: float, pts, *, ctx: Context | None = None) -> EllipseTree:
    return EllipseTree(sf, pts, ctx=ctx)
