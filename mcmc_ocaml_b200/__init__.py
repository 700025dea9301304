"""mcmc_ocaml_b200 -- B200-native sampling-and-evidence path behind the API
shape of farr/mcmc-ocaml (Mcmc, Kd_tree, Interpolate_pdf, Evidence, Stats,
Nested).  All compute runs in hand-written sm_100a CUDA kernels inside
libmcmcgpu.so (C ABI: include/mcmc_gpu.h); this package is the host-side
mirror of the reference's module interface.  There is no CPU fallback.
"""
from . import _abi, plugins
from ._abi import Failure, InvalidArgument
from .context import Context, default_context

__all__ = ["Context", "default_context", "plugins", "Failure", "InvalidArgument"]
