"""Ellipse on one B200: enclosing ellipse and ellipse tree of N x D points, device resident, next
to the numpy restatement of ellipse.ml on a bounded sample.  Prints one JSON line.
usage: python tools/bench_ellipse.py [--n 4000000] [--d 5] [--reps 3] [--cpu-n 200000]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_ocaml_b200 import Context, ellipse


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4_000_000); ap.add_argument("--d", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3); ap.add_argument("--cpu-n", type=int, default=200_000)
    a = ap.parse_args()
    ctx = Context(0, 1)
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    scale = torch.linspace(0.5, 2.0, a.d, dtype=torch.float64, device="cuda")
    x = torch.empty((a.n, a.d), dtype=torch.float64, device="cuda").normal_(0.0, 1.0, generator=g) * scale
    torch.cuda.synchronize()
    out = dict(N=a.n, D=a.d)
    best = 1e9
    for _ in range(a.reps + 1):
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        t = ellipse.EllipseTree(2.0, device_ptr=x.data_ptr(), n=a.n, dim=a.d, ctx=ctx)
        dt = time.perf_counter() - t0
        out.update(tree_nodes=t.nnodes, tree_levels=t.nlevels, tree_launches=ctx.launch_count - l0)
        t.close()
        best = min(best, dt)
    out.update(tree_s=best, tree_points_per_s=a.n / best,
               tree_algorithmic_bytes=int(out["tree_levels"]) * a.n * (4 * 8 * a.d + 5 * 8 * a.d // 4),
               note="bytes: per level 3 read passes (mean, covariance, range) + the flag pass + read and write of the rows in the scatter")
    # CPU: the numpy restatement (oracle) on a bounded sample of the same distribution
    from oracle import ellipse_np as E
    xs = x[:a.cpu_n].cpu().numpy()
    t0 = time.perf_counter(); ot = E.ellipse_tree(2.0, xs); cdt = time.perf_counter() - t0
    out.update(cpu_baseline=dict(kind="port", cores=1, sample=f"numpy restatement of ellipse.ml (oracle/ellipse_np.py) on {a.cpu_n} x {a.d} points of the same distribution",
                                 seconds=cdt, points_per_s=a.cpu_n / cdt, nodes=len(E.flatten(ot)["ids"])))
    print(json.dumps(out))


main()
