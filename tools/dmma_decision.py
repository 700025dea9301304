#!/usr/bin/env python
"""FP64 tensor cores for the quadratic form of a high-dimensional correlated Gaussian: measurements behind the
keep-or-drop decision (north star: "tensor cores only for the batched (x - mu)^T Sigma^-1 (x - mu)").
Prints one JSON document: FP64 FMA and DMMA peaks of this GPU, and for D in {32, 64} the time of 2^22 evaluations by
(0) the sampler plugin's FMA order, one thread per point, and (1) mma.sync.m8n8k4.f64, one warp per 32 points;
agreement of the two (relative) and of variant 0 with mg_logfn_eval's MG_FN_GAUSS_CORR (must be identical)."""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(M=1 << 22, reps=5, out=None):
    import numpy as np
    import torch

    from mcmc_ocaml_b200 import Context, _abi, plugins as P
    ctx = Context(0, 1)
    lib = ctx.lib
    doc = {"M": M}
    v = C.c_double()
    lib.mg_measure_fp64_tflops(ctx.h, 3, C.byref(v)); doc["fp64_fma_tflops"] = v.value
    lib.mg_measure_dmma_tflops(ctx.h, 3, C.byref(v)); doc["fp64_dmma_tflops"] = v.value
    lib.mg_debug_quadform.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _abi.c_double_p, _abi.c_double_p, C.c_double, C.c_void_p,
                                      C.c_int64, C.c_void_p, C.c_int32, C.POINTER(C.c_double)]
    doc["cases"] = []
    for D in (32, 64):
        mu = np.arange(D) / 10.0
        cov = 0.7 ** np.abs(np.subtract.outer(np.arange(D), np.arange(D)))
        like = P.gauss_corr(mu, cov)
        par = _abi.as_f64(like.params)
        Lp, logc = par[D:D + D * (D + 1) // 2].copy(), float(par[D + D * (D + 1) // 2])
        g = torch.Generator(device="cuda"); g.manual_seed(D)
        x = torch.randn((M, D), dtype=torch.float64, device="cuda", generator=g) + torch.as_tensor(mu, device="cuda")
        xt = x.t().contiguous()          # [D][M], the sampler's state layout
        o0 = torch.empty(M, dtype=torch.float64, device="cuda"); o1 = torch.empty_like(o0)
        torch.cuda.synchronize()
        ms0, ms1 = C.c_double(), C.c_double()
        ctx.check(lib.mg_debug_quadform(ctx.h, 0, D, _abi.ptr(mu), _abi.ptr(Lp), logc, xt.data_ptr(), M, o0.data_ptr(), reps, C.byref(ms0)))
        ctx.check(lib.mg_debug_quadform(ctx.h, 1, D, _abi.ptr(mu), _abi.ptr(Lp), logc, xt.data_ptr(), M, o1.data_ptr(), reps, C.byref(ms1)))
        ctx.sync()
        # variant 0 against the sampler's plugin on a sample of the points
        xs = x[:20000].cpu().numpy()
        ref = np.empty(20000)
        s = like.spec()
        ctx.check(lib.mg_logfn_eval(ctx.h, C.byref(s), _abi.ptr(xs), C.c_int64(20000), _abi.ptr(ref)))
        a0, a1 = o0.cpu().numpy(), o1.cpu().numpy()
        flops = (D * (D + 1) + 2 * D + 2) * M          # D(D+1)/2 FMAs for L z, D for the squares, D subtractions
        doc["cases"].append({"D": D, "fma_ms": ms0.value, "dmma_ms": ms1.value, "speedup_dmma": ms0.value / ms1.value,
                             "fma_tflops": flops / (ms0.value * 1e-3) / 1e12, "dmma_useful_tflops": flops / (ms1.value * 1e-3) / 1e12,
                             "fma_identical_to_plugin": bool(np.array_equal(a0[:20000], ref)),
                             "max_rel_diff_dmma_vs_fma": float(np.max(np.abs(a1 - a0) / np.maximum(1.0, np.abs(a0))))})
    s_ = json.dumps(doc, indent=1)
    print(s_)
    if out:
        with open(out, "w") as f:
            f.write(s_ + "\n")
    return doc


if __name__ == "__main__":
    run(out=sys.argv[1] if len(sys.argv) > 1 else None)
