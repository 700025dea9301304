#!/usr/bin/env python
"""BASELINE.json config 2 away from the headline point: nskip = 10 (float64-compute bound instead of HBM bound) and
other dimensions, device resident, kernel time from the library's own CUDA events.  One JSON line per variant."""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch

    from mcmc_ocaml_b200 import Context, _abi, plugins as P
    ctx = Context(0, 0x5EED0001)
    Cn, T = 65536, 10000
    for D, nskip in [(10, 1), (10, 10), (10, 100), (4, 1), (20, 1), (32, 10)]:
        mu = np.arange(D) / 10.0
        cov = 0.7 ** np.abs(np.subtract.outer(np.arange(D), np.arange(D)))
        like, prior, prop = P.gauss_corr(mu, cov), P.zero(D), P.box_proposal(np.full(D, 0.5 if D <= 10 else 0.3))
        F = D + 2
        n = T // nskip + 1
        if n * F * Cn * 8 > 120e9:
            n = int(120e9 // (F * Cn * 8)); 
        steps = (n - 1) * nskip
        blk = torch.empty((n, F, Cn), dtype=torch.float64, device="cuda")
        state = torch.empty((F, Cn), dtype=torch.float64, device="cuda")
        state[:D] = torch.as_tensor(mu, device="cuda")[:, None]
        acc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        cfg = _abi.mg_mcmc_cfg(Cn, D, 0, 0, nskip, n, 0, 1, 0)
        ls, ps, js = like.spec(), prior.spec(), prop.spec()
        best = 1e30
        for _ in range(3):
            ctx.check(ctx.lib.mg_mcmc_array_dev(ctx.h, C.byref(ls), C.byref(ps), C.byref(js), C.byref(cfg),
                                                C.c_void_p(state.data_ptr()), C.c_void_p(blk.data_ptr()),
                                                C.c_void_p(acc.data_ptr())))
            ctx.sync()
            best = min(best, ctx.last_kernel_ms)
        flops = D * D + 8 * D + 3
        print(json.dumps({"dim": D, "nskip": nskip, "chains": Cn, "steps": steps, "kernel_ms": best,
                          "chain_steps_per_s": Cn * steps / (best * 1e-3),
                          "write_GBps": 8 * F * Cn * (n - 1) / (best * 1e-3) / 1e9,
                          "fp64_tflops": flops * Cn * steps / (best * 1e-3) / 1e12}))
        del blk, state, acc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
