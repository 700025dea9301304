"""Times truncated kd-tree builds (min_split far above the subtree size) on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_ocaml_b200 import Context, kd_tree
N, D = int(sys.argv[1]), int(sys.argv[2])
ctx = Context(0, 1)
g = torch.Generator(device="cuda"); g.manual_seed(12345)
x = torch.empty((N, D), dtype=torch.float64, device="cuda").normal_(0.5, 0.05, generator=g)
torch.cuda.synchronize()
for ms in [int(a) for a in sys.argv[3:]]:
    ts = []
    for _ in range(4):
        t = time.perf_counter()
        tr = kd_tree.KdTree.from_device(x.data_ptr(), N, D, np.zeros(D), np.ones(D), min_split=ms, ctx=ctx)
        ctx.sync(); ts.append(time.perf_counter() - t)
        nn, nl = tr.nnodes, tr.nlevels
        tr.close()
    print(ms, nn, nl, ["%.2f ms" % (1e3 * v) for v in ts], flush=True)
