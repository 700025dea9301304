"""Where the host-buffer Lebesgue call spends its time: H2D alone, device call alone, the host call (per repetition)."""
import ctypes as C, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_ocaml_b200 import Context, evidence
N, D = 10_000_000, 20
ctx = Context(0, 1)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(12345)
x = torch.empty((N, D), dtype=torch.float64, device=dev).normal_(0.5, 0.05, generator=g)
ll = (-0.91893853320467274178 - math.log(0.05) - 0.5 * ((x - 0.5) / 0.05) ** 2).sum(1)
lp = torch.zeros(N, dtype=torch.float64, device=dev)
xh = torch.empty((N, D), dtype=torch.float64).pin_memory(); xh.copy_(x)
llh = torch.empty(N, dtype=torch.float64).pin_memory(); llh.copy_(ll)
lph = torch.zeros(N, dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
for _ in range(3):
    t = time.perf_counter(); x.copy_(xh, non_blocking=True); torch.cuda.synchronize(); print("torch H2D 1.6 GB: %.2f ms (%.1f GB/s)" % (1e3 * (time.perf_counter() - t), 1.6 / (time.perf_counter() - t)))
for _ in range(3):
    t = time.perf_counter(); z = evidence.evidence_lebesgue_dev(x.data_ptr(), ll.data_ptr(), lp.data_ptr(), N, D, ctx=ctx); print("device call: %.2f ms" % (1e3 * (time.perf_counter() - t)))
zc = C.c_double()
for _ in range(6):
    t = time.perf_counter()
    ctx.check(ctx.lib.mg_evidence_lebesgue(ctx.h, C.c_void_p(xh.data_ptr()), C.c_void_p(llh.data_ptr()), C.c_void_p(lph.data_ptr()),
                                           C.c_int64(N), C.c_int32(D), C.c_int32(64), C.c_double(0.1), C.byref(zc)))
    print("host call: %.2f ms  Z = %r (device %r)" % (1e3 * (time.perf_counter() - t), zc.value, z))
