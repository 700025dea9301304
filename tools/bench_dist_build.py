#!/usr/bin/env python
"""Distributed kd-tree build (mg_kdtree_build_distributed) against the single-GPU build: torchrun, one rank per GPU.
Every rank holds the same N x D rows (generated on the device from one seed).  Prints one JSON line on rank 0.
usage: torchrun ... tools/bench_dist_build.py [--points 10000000] [--dim 20] [--min-split 2] [--reps 4] [--check]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", dest="n", type=int, default=10_000_000); ap.add_argument("--dim", dest="d", type=int, default=20)
    ap.add_argument("--min-split", type=int, default=2); ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import numpy as np, torch, torch.distributed as dist
    from mcmc_ocaml_b200 import Context, comm as CM, kd_tree
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("gloo")
    ctx = Context(local, 1)
    comm = CM.Comm.from_torch(ctx)
    g = torch.Generator(device=dev); g.manual_seed(12345)
    x = torch.empty((a.n, a.d), dtype=torch.float64, device=dev).normal_(0.5, 0.05, generator=g)
    lo, hi = np.zeros(a.d), np.ones(a.d)
    torch.cuda.synchronize()
    td_times, ts_times = [], []
    td = ts = None
    for rep in range(a.reps):
        if td is not None: td.close()
        comm.barrier(); torch.cuda.synchronize()
        t = time.perf_counter()
        td = comm.build_tree(x.data_ptr(), a.n, a.d, lo, hi, min_split=a.min_split)
        ctx.sync(); td_times.append(time.perf_counter() - t)
    for rep in range(a.reps):
        if ts is not None: ts.close()
        comm.barrier(); torch.cuda.synchronize()
        t = time.perf_counter()
        ts = kd_tree.KdTree.from_device(x.data_ptr(), a.n, a.d, lo, hi, min_split=a.min_split, ctx=ctx)
        ctx.sync(); ts_times.append(time.perf_counter() - t)
    same = None
    if a.check:
        ad, as_ = td.export(), ts.export()
        same = bool(all(np.array_equal(ad[k], as_[k]) for k in as_) and td.nnodes == ts.nnodes and td.nlevels == ts.nlevels)
        same = bool(comm.allgather(np.array([1.0 if same else 0.0])).min() == 1.0)
    worst = comm.allgather(np.array([min(td_times[1:]), min(ts_times[1:])]))
    if rank == 0:
        d_s, s_s = float(worst[:, 0].max()), float(worst[:, 1].max())
        print(json.dumps(dict(world=world, N=a.n, D=a.d, min_split=a.min_split, nnodes=int(td.nnodes), nlevels=int(td.nlevels),
                              distributed_s=d_s, single_gpu_s=s_s, speedup=s_s / d_s, identical_on_every_rank=same,
                              points_per_s_distributed=a.n / d_s, note="max over ranks of each rank's best repetition; the distributed build returns the whole tree on every rank")))
    comm.barrier(); comm.close()
    if world > 1:
        dist.destroy_process_group()


main()
