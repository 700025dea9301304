"""Two kd-tree builds at N x D (min_split from argv[3], default 2), device resident (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_ocaml_b200 import Context, kd_tree
N, D = int(sys.argv[1]), int(sys.argv[2])
ms = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = Context(0, 1)
g = torch.Generator(device="cuda"); g.manual_seed(12345)
x = torch.empty((N, D), dtype=torch.float64, device="cuda").normal_(0.5, 0.05, generator=g)
torch.cuda.synchronize()
import time
for _ in range(3):
    t = time.perf_counter()
    tr = kd_tree.KdTree.from_device(x.data_ptr(), N, D, np.zeros(D), np.ones(D), min_split=ms, ctx=ctx)
    dt = time.perf_counter() - t
    print(tr.nnodes, tr.nlevels, dt)
    tr.close()
