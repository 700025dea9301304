"""Context: one GPU, one Philox key, one CUDA stream (mg_ctx in the C ABI)."""
from __future__ import annotations

import ctypes as C

from . import _abi


class Context:
    """Owns an ``mg_ctx``.  Single caller at a time, like the reference
    (global counters mcmc.ml:27-28, global ``Random`` state)."""

    def __init__(self, device: int = 0, seed: int = 0):
        self.lib = _abi.load_library()
        h = C.c_void_p()
        rc = self.lib.mg_ctx_create(int(device), C.c_uint64(seed & (2**64 - 1)), C.byref(h))
        if rc != _abi.MG_OK or not h:
            raise _abi.Failure(f"cuda: cannot create a context on device {device} (status {rc}); "
                               "the GPU path has no CPU fallback")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.mg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- error mapping: EINVAL -> Invalid_argument, else Failure
    def check(self, rc: int):
        if rc == _abi.MG_OK:
            return
        msg = self.lib.mg_last_error(self.h).decode("utf-8", "replace")
        if rc == _abi.MG_EINVAL:
            raise _abi.InvalidArgument(msg)
        raise _abi.Failure(msg)

    def set_seed(self, seed: int):
        """``Random.init seed``."""
        self.check(self.lib.mg_ctx_set_seed(self.h, C.c_uint64(seed & (2**64 - 1))))

    @property
    def epoch(self) -> int:
        return int(self.lib.mg_ctx_get_epoch(self.h))

    @epoch.setter
    def epoch(self, e: int):
        self.check(self.lib.mg_ctx_set_epoch(self.h, C.c_uint64(e)))

    def set_stream(self, cuda_stream: int | None):
        self.check(self.lib.mg_ctx_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self.check(self.lib.mg_ctx_sync(self.h))

    def trim_pool(self):
        """Hand the library's cached temporaries (stream-ordered pool) back to the driver."""
        self.check(self.lib.mg_ctx_trim_pool(self.h))

    def reserve_pool(self, gigabytes: float = 12.0):
        """Make the pool hold at least this much in one piece (``mg_ctx_reserve_pool``; done once by ``mg_ctx_create``,
        to be repeated after ``trim_pool``)."""
        self.check(self.lib.mg_ctx_reserve_pool(self.h, C.c_int64(int(gigabytes * (1 << 30)))))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mg_ctx_launch_count(self.h))

    @property
    def last_kernel_ms(self) -> float:
        return float(self.lib.mg_ctx_last_kernel_ms(self.h))

    # Mcmc.reset_counters / Mcmc.get_counters (mcmc.ml:30-35)
    def reset_counters(self):
        self.check(self.lib.mg_reset_counters(self.h))

    def get_counters(self) -> tuple[int, int]:
        a, r = C.c_int64(), C.c_int64()
        self.check(self.lib.mg_get_counters(self.h, C.byref(a), C.byref(r)))
        return int(a.value), int(r.value)


_default: Context | None = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(0, 0)
    return _default
