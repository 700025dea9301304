"""Host-side mirror of ``Kd_tree.Make`` (kd_tree.ml:31-60) for the GPU path:
flat node arrays in HBM instead of ``Cell of o list * ... * tree * tree``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .context import Context, default_context


class KdTree:
    """``tree_of_objects objs low high`` (kd_tree.ml:155-175), built on the GPU."""

    def __init__(self, pts, low, high, *, min_split: int = 2, ctx: Context | None = None, _handle=None):
        self.ctx = ctx or default_context()
        self._box = None
        if _handle is not None:
            self.h = _handle
        else:
            pts = _abi.as_f64(pts)
            if pts.ndim == 1:
                pts = pts.reshape(-1, 1)
            low, high = _abi.as_f64(low), _abi.as_f64(high)
            if pts.ndim != 2 or low.size != pts.shape[1] or high.size != pts.shape[1]:
                raise _abi.InvalidArgument("tree_of_objects: bounds do not match the points")
            h = C.c_void_p()
            self.ctx.check(self.ctx.lib.mg_kdtree_build(self.ctx.h, _abi.ptr(pts), C.c_int64(pts.shape[0]),
                                                        C.c_int32(pts.shape[1]), _abi.ptr(low), _abi.ptr(high),
                                                        C.c_int32(min_split), C.byref(h)))
            self.h = h
            self._box = (low.copy(), high.copy())
        info = self.info()
        self.N, self.D, self.nnodes, self.nlevels = info["npoints"], info["dim"], info["nnodes"], info["nlevels"]

    @classmethod
    def from_device(cls, pts_ptr: int, N: int, D: int, low, high, *, min_split: int = 2, ctx: Context | None = None):
        """Build from a device array float64 [N][D] (e.g. a torch tensor's data_ptr())."""
        ctx = ctx or default_context()
        low, high = _abi.as_f64(low), _abi.as_f64(high)
        h = C.c_void_p()
        ctx.check(ctx.lib.mg_kdtree_build_dev(ctx.h, C.c_void_p(pts_ptr), C.c_int64(N), C.c_int32(D), _abi.ptr(low),
                                              _abi.ptr(high), C.c_int32(min_split), C.byref(h)))
        t = cls(None, None, None, ctx=ctx, _handle=h)
        t._box = (low.copy(), high.copy())
        return t

    @classmethod
    def from_blob(cls, blob_ptr: int, nbytes: int, *, ctx: Context | None = None):
        """Rebuild a tree from a serialised device blob (after an NCCL broadcast)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.mg_kdtree_from_blob_dev(ctx.h, C.c_void_p(blob_ptr), C.c_int64(nbytes), C.byref(h)))
        return cls(None, None, None, ctx=ctx, _handle=h)

    def blob(self) -> tuple[int, int]:
        """(device pointer, nbytes) of the contiguous serialised tree (borrowed)."""
        n, p = C.c_int64(), C.c_void_p()
        self.ctx.check(self.ctx.lib.mg_kdtree_blob_size(self.h, C.byref(n)))
        self.ctx.check(self.ctx.lib.mg_kdtree_blob_dev(self.h, C.byref(p)))
        return int(p.value), int(n.value)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mg_kdtree_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        n, d, nn, nl = C.c_int64(), C.c_int32(), C.c_int64(), C.c_int32()
        self.ctx.check(self.ctx.lib.mg_kdtree_info(self.h, C.byref(n), C.byref(d), C.byref(nn), C.byref(nl)))
        return dict(npoints=n.value, dim=d.value, nnodes=nn.value, nlevels=nl.value)

    def export(self):
        """Flat arrays for bit-exact comparison (see mg_kdtree_export)."""
        nn = self.nnodes
        sd = np.empty(nn, np.int32); sv = np.empty(nn); left = np.empty(nn, np.int32)
        b = np.empty(nn, np.int32); e = np.empty(nn, np.int32); perm = np.empty(self.N, np.int32)
        self.ctx.check(self.ctx.lib.mg_kdtree_export(self.h, _abi.ptr(sd, _abi.c_int32_p), _abi.ptr(sv),
                                                     _abi.ptr(left, _abi.c_int32_p), _abi.ptr(b, _abi.c_int32_p),
                                                     _abi.ptr(e, _abi.c_int32_p), _abi.ptr(perm, _abi.c_int32_p)))
        return dict(split_dim=sd, split_val=sv, left=left, begin=b, end=e, perm=perm)

    def volume(self) -> float:
        """``Kd_tree.volume tree`` (kd_tree.ml:184-186): the volume of the root cell's (caller supplied) box."""
        if self._box is None:
            raise _abi.InvalidArgument("volume: the root box of a tree rebuilt from a blob is held on the device only")
        return bounds_volume(*self._box)

    def depth(self) -> int:
        """``depth`` of test/kd_tree_test.ml:66-71 (number of levels)."""
        return self.nlevels


def tree_of_objects(objs, low, high, **kw) -> KdTree:
    return KdTree(objs, low, high, **kw)


def bounds_of_objects(objs):
    """``Kd_tree.bounds_of_objects`` (kd_tree.ml:96-110)."""
    objs = _abi.as_f64(objs)
    if objs.size == 0:
        raise _abi.InvalidArgument("bounds_of_objects: no objects")
    return objs.min(axis=0), objs.max(axis=0)


def bounds_volume(low, high) -> float:
    """``Kd_tree.bounds_volume`` (kd_tree.ml:177-182): left-to-right product."""
    low, high = _abi.as_f64(low), _abi.as_f64(high)
    return float(_abi.load_library().mg_bounds_volume(_abi.ptr(low), _abi.ptr(high), C.c_int32(low.size)))
