#!/usr/bin/env python
"""Multi-GPU plumbing check (torchrun, one rank per GPU, NCCL):
  1. rank 0 builds a kd-tree, the serialised blob is broadcast to every rank
     (one NCCL broadcast over NVLink), each rank evaluates Interpolate_pdf
     densities on ITS shard of the queries, results are all-gathered and
     compared with rank 0's own evaluation of all queries (bit-exact);
  2. every rank runs its shard of an MH ensemble (global chain ids), the
     per-rank block statistics are all-gathered and combined; the pooled mean /
     std must equal the single-rank run over all chains (1e-12).
Prints one JSON line on rank 0."""
from __future__ import annotations

import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    from mcmc_ocaml_b200 import Context, distributed as D, interpolate_pdf, kd_tree, mcmc, plugins as P, stats
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local, 2024)
    out = {"world": world}
    # ---- 1. tree broadcast + sharded queries ---------------------------------
    N, Dm, M = 2_000_000, 8, 1_000_000
    rng = np.random.default_rng(7)
    q = rng.random((M, Dm)) * 0.4 + 0.3
    tree = None
    if rank == 0:
        pts = rng.normal(0.5, 0.08, (N, Dm)).clip(0, 1)
        tree = kd_tree.KdTree(pts, np.zeros(Dm), np.ones(Dm), ctx=ctx)
    torch.cuda.synchronize(); t = time.perf_counter()
    tree = D.broadcast_tree(tree, 0, ctx=ctx)
    torch.cuda.synchronize()
    out["tree_broadcast_s"] = time.perf_counter() - t
    out["tree_blob_MB"] = tree.blob()[1] / 1e6
    b, e = D.shard_range(M, rank, world)
    interp = interpolate_pdf.InterpPdf(None, None, None, tree=tree)
    mine = interp.jump_prob(q[b:e])
    full = torch.zeros(M, dtype=torch.float64, device=dev)
    full[b:e] = torch.as_tensor(mine, device=dev)
    if world > 1:
        dist.all_reduce(full)          # disjoint shards: a sum assembles the whole vector
    if rank == 0:
        want = interp.jump_prob(q)
        out["sharded_density_bit_exact"] = bool(np.array_equal(full.cpu().numpy(), want))
    # ---- 2. sharded MH ensemble + gathered statistics ------------------------
    Dd, C, n = 10, 4096, 200
    mu = np.arange(Dd) / 10.0
    cov = 0.7 ** np.abs(np.subtract.outer(np.arange(Dd), np.arange(Dd)))
    like, prior, prop = P.gauss_corr(mu, cov), P.zero(Dd), P.box_proposal(np.full(Dd, 0.5))
    cb, ce = D.shard_range(C, rank, world)
    ctx.set_seed(99)
    s = mcmc.mcmc_array(n, like, prior, prop, mu, nchains=ce - cb, chain_offset=cb, nbin=20, ctx=ctx)
    pooled = s.block.transpose(0, 2, 1).reshape(-1, Dd + 2)
    res = D.gather_ensemble_stats(pooled.shape[0], pooled.mean(0), pooled.std(0, ddof=1), int(s.accept.sum()), int(s.reject.sum()), device=dev)
    # ---- 3. harmonic-mean evidence of the pooled chains from per-rank shards ---
    z_sharded = D.harmonic_mean_sharded(pooled[:, Dd], ctx=ctx, device=dev)
    if rank == 0:
        ctx.set_seed(99)
        ref = mcmc.mcmc_array(n, like, prior, prop, mu, nchains=C, chain_offset=0, nbin=20, ctx=ctx)
        from mcmc_ocaml_b200 import evidence
        z_all = evidence.evidence_harmonic_mean(ll=ref.block[:, Dd, :].reshape(-1), ctx=ctx)
        out["harmonic_sharded_rel_err"] = float(abs(z_sharded - z_all) / abs(z_all))
        rp = ref.block.transpose(0, 2, 1).reshape(-1, Dd + 2)
        out["ensemble_mean_err"] = float(np.max(np.abs(res["mean"] - rp.mean(0))))
        out["ensemble_std_err"] = float(np.max(np.abs(res["std"] - rp.std(0, ddof=1))))
        out["accept_equal"] = bool(res["accept"] == int(ref.accept.sum()))
        out["shard_chains_bit_exact"] = bool(np.array_equal(ref.block[:, :, cb:ce], s.block))
        out["ok"] = bool(out.get("sharded_density_bit_exact") and out["accept_equal"] and out["shard_chains_bit_exact"]
                         and out["ensemble_mean_err"] < 1e-12 and out["ensemble_std_err"] < 1e-12
                         and out["harmonic_sharded_rel_err"] < 1e-12)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
