"""Host-side mirror of ``Read_write`` (read_write.mli:31-65): the text format
the reference's tools use to move samples between runs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


def _lib():
    lib = _abi.load_library()
    lib.mg_free_host.restype = None
    lib.mg_free_host.argtypes = [C.c_void_p]
    return lib


def _check(rc, what):
    if rc == _abi.MG_OK:
        return
    if rc == _abi.MG_EINVAL:
        raise _abi.InvalidArgument(f"{what}: malformed sample file")
    raise _abi.Failure(f"{what}: cannot read / write / parse")


def write(path: str, rows, dim: int | None = None, *, lossless: bool = False) -> None:
    """``Read_write.write to_coords chan samples``; rows [n][D+2] (value, log_likelihood, log_prior).
    ``lossless`` writes "%.17g" instead of the reference's "%g"."""
    rows = _abi.as_f64(rows)
    if rows.ndim != 2 or rows.shape[1] < 2:
        raise _abi.InvalidArgument("write: rows must be [n][D+2]")
    D = rows.shape[1] - 2 if dim is None else dim
    _check(_lib().mg_write_samples(path.encode(), _abi.ptr(rows), C.c_int64(rows.shape[0]), C.c_int32(D),
                                   C.c_int32(17 if lossless else 0)), "write")


def read(path: str) -> np.ndarray:
    """``Read_write.read from_coords chan``: rows [n][D+2]."""
    lib = _lib()
    p, n, D = C.POINTER(C.c_double)(), C.c_int64(), C.c_int32()
    _check(lib.mg_read_samples(path.encode(), C.byref(p), C.byref(n), C.byref(D)), "read")
    try:
        out = np.ctypeslib.as_array(p, shape=(n.value, D.value + 2)).copy() if n.value else np.empty((0, 2))
    finally:
        lib.mg_free_host(p)
    return out


def write_nested(path: str, log_ev: float, log_dev: float, rows, log_weights, *, lossless: bool = False) -> None:
    """``Read_write.write_nested``; rows [n][D+2], log_weights [n]."""
    rows, lw = _abi.as_f64(rows), _abi.as_f64(log_weights)
    _check(_lib().mg_write_nested(path.encode(), C.c_double(log_ev), C.c_double(log_dev), _abi.ptr(rows), _abi.ptr(lw),
                                  C.c_int64(rows.shape[0]), C.c_int32(rows.shape[1] - 2), C.c_int32(17 if lossless else 0)),
           "write_nested")


def read_nested(path: str):
    """``Read_write.read_nested``: (log_ev, log_dev, rows [n][D+2], log_weights [n])."""
    lib = _lib()
    p, w, n, D = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.c_int64(), C.c_int32()
    lev, ldev = C.c_double(), C.c_double()
    _check(lib.mg_read_nested(path.encode(), C.byref(lev), C.byref(ldev), C.byref(p), C.byref(w), C.byref(n), C.byref(D)),
           "read_nested")
    try:
        rows = np.ctypeslib.as_array(p, shape=(n.value, D.value + 2)).copy() if n.value else np.empty((0, 2))
        lw = np.ctypeslib.as_array(w, shape=(n.value,)).copy() if n.value else np.empty(0)
    finally:
        lib.mg_free_host(p); lib.mg_free_host(w)
    return lev.value, ldev.value, rows, lw
