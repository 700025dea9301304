"""One Lebesgue evidence call at N x D, device resident (for ncu launch lists)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcmc_ocaml_b200 import Context, evidence
N, D = int(sys.argv[1]), int(sys.argv[2])
ctx = Context(0, 1)
g = torch.Generator(device="cuda"); g.manual_seed(12345)
x = torch.empty((N, D), dtype=torch.float64, device="cuda").normal_(0.5, 0.05, generator=g)
ll = (-0.91893853320467274178 - math.log(0.05) - 0.5 * ((x - 0.5) / 0.05) ** 2).sum(1)
lp = torch.zeros(N, dtype=torch.float64, device="cuda")
for _ in range(2):
    z = evidence.evidence_lebesgue_dev(x.data_ptr(), ll.data_ptr(), lp.data_ptr(), N, D, ctx=ctx)
print(z)
