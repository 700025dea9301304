"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from mcmc_ocaml_b200 import distributed as D


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_combine_moments_matches_pooled():
    rng = np.random.default_rng(0)
    parts = [rng.normal(3, 2, (n, 4)) for n in (1000, 1, 2500, 333)]
    n, mu, sd = D.combine_moments([len(p) for p in parts], [p.mean(0) for p in parts],
                                  [((p - p.mean(0)) ** 2).sum(0) for p in parts])
    allx = np.concatenate(parts)
    assert n == len(allx)
    np.testing.assert_allclose(mu, allx.mean(0), rtol=1e-13)
    np.testing.assert_allclose(sd, allx.std(0, ddof=1), rtol=1e-13)


def test_combine_harmonic_matches_pooled():
    rng = np.random.default_rng(2)
    ll = rng.normal(-3.0, 1.0, 5000)
    pooled = ll.size / np.sum(1.0 / np.exp(ll))
    parts = [ll[:1234], ll[1234:1235], ll[1235:]]
    z = D.combine_harmonic([p.size for p in parts], [p.size / np.sum(1.0 / np.exp(p)) for p in parts])
    assert abs(z - pooled) <= 1e-13 * pooled


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    b, e = D.shard_range(1001, rank, world)
    x = np.random.default_rng(5).normal(1.0, 3.0, (1001, 3))[b:e]       # this rank's shard of a common data set
    res = D.gather_ensemble_stats(len(x), x.mean(0), x.std(0, ddof=1), accept=10 * (rank + 1), reject=5)
    g = D.all_gather_array(np.array([float(rank), rng.random()]))
    na, nb, ratio = D.combine_model_counts([100 * (rank + 1), 50])
    assert (na, nb, ratio) == (300, 100, 3.0)
    # the 128-byte NCCL id of the C-ABI communicator travels from rank 0 to the others through the host program's
    # own channel (mcmc_ocaml_b200/comm.py): here a gloo broadcast and a file
    from mcmc_ocaml_b200 import comm as CM
    make = lambda: bytes((7 * i + 3) % 256 for i in range(CM.ID_BYTES))
    assert CM.exchange_id_torch(make if rank == 0 else (lambda: b""), rank) == make()
    path = os.path.join(os.environ["MG_TEST_TMP"], "nccl_id.bin")
    assert CM.exchange_id_file(make if rank == 0 else (lambda: b""), rank, path) == make()
    q.put((rank, res["n"], res["mean"], res["std"], res["accept"], res["reject"], g[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_and_combine(tmp_path, monkeypatch):
    monkeypatch.setenv("MG_TEST_TMP", str(tmp_path))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    x = np.random.default_rng(5).normal(1.0, 3.0, (1001, 3))
    for rank, n, mu, sd, acc, rej, ranks in out:
        assert n == 1001 and acc == 30 and rej == 10 and ranks == [0.0, 1.0]
        np.testing.assert_allclose(mu, x.mean(0), rtol=1e-12)
        np.testing.assert_allclose(sd, x.std(0, ddof=1), rtol=1e-12)


@pytest.mark.parametrize("n,d,ranks,min_split", [(20000, 3, 2, 2), (30011, 5, 4, 2), (16384, 2, 8, 2), (50000, 4, 4, 64)])
def test_distributed_build_rule_reproduces_the_whole_tree(n, d, ranks, min_split):
    """The numbering rule of mg_kdtree_build_distributed (D.graft_subtrees mirrors csrc/comm.cu) on oracle trees: the
    top truncated at (N >> k) + 3, one complete subtree per leaf built from the leaf's rows in the top's order, grafted
    -> every array of the oracle's own whole tree, bit for bit."""
    from oracle import oracle as og
    rng = np.random.default_rng(n + ranks)
    pts = rng.normal(0.5, 0.1, (n, d))
    lo, hi = np.zeros(d) - 5, np.ones(d) + 5
    k = ranks.bit_length() - 1
    whole = og.Tree(pts, lo, hi, min_split=min_split).export()
    top = og.Tree(pts, lo, hi, min_split=(n >> k) + 3).export()
    assert len(top["left"]) == 2 * ranks - 1 and np.all(top["left"][:ranks - 1] == 2 * np.arange(ranks - 1) + 1)
    subs = []
    for r in range(ranks):
        b, e = top["begin"][ranks - 1 + r], top["end"][ranks - 1 + r]
        rows = np.ascontiguousarray(pts[top["perm"][b:e]])
        subs.append(og.Tree(rows, lo, hi, min_split=min_split).export())
    got = D.graft_subtrees(top, subs)
    for key in whole:
        assert np.array_equal(got[key], whole[key]), key
    assert D.tree_levels(whole["left"])[-1] == len(whole["left"])
