"""Small C-ABI entry points: plugin evaluation, remove_repeat_samples, device helpers."""
import ctypes as C

import numpy as np
import pytest

from mcmc_ocaml_b200 import Failure, _abi, mcmc, plugins as P

pytestmark = pytest.mark.gpu


def eval_gpu(ctx, fn, x):
    x = _abi.as_f64(x).reshape(-1, fn.dim)
    out = np.empty(x.shape[0])
    s = fn.spec()
    ctx.check(ctx.lib.mg_logfn_eval(ctx.h, C.byref(s), _abi.ptr(x), C.c_int64(x.shape[0]), _abi.ptr(out)))
    return out


def test_every_builtin_logfn_matches_oracle(ctx, og):
    from tests.golden.gc_data import DATA
    rng = np.random.default_rng(0)
    D = 5
    mu = rng.random(D); sig = 0.5 + rng.random(D)
    A = rng.normal(size=(D, D)); cov = A @ A.T + D * np.eye(D)
    fns = [P.zero(D), P.const(D, -1.25), P.box(np.zeros(D), np.ones(D), -0.3), P.box(np.zeros(D), np.ones(D), 0.0, closed=False),
           P.gauss_diag(mu, sig), P.gauss_corr(mu, cov), P.shell(mu, 2.0, 0.1),
           P.gauss_mix(rng.random((3, D)), sig), P.gauss_diag(mu, sig).scaled(0.75)]
    x = rng.normal(0.5, 1.0, (2000, D))
    x[:10] = np.clip(x[:10], 0, 1)
    for fn in fns:
        got, want = eval_gpu(ctx, fn, x), og.logfn_eval(fn, x)
        exact = fn.kind in (_abi.FN_ZERO, _abi.FN_CONST, _abi.FN_BOX_CLOSED, _abi.FN_BOX_OPEN, _abi.FN_GAUSS_CORR)
        if exact:
            assert np.array_equal(got, want), fn.kind
        else:
            np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13, err_msg=str(fn.kind))
    x2 = np.stack([rng.uniform(-1, 1, 500), rng.uniform(0.5, 1.5, 500)], axis=1)
    for fn in (P.gauss_data(DATA), P.cauchy_data(DATA)):
        np.testing.assert_allclose(eval_gpu(ctx, fn, x2), og.logfn_eval(fn, x2), rtol=1e-13)


def test_remove_repeat_samples(ctx, og):
    """Mcmc.remove_repeat_samples (mcmc.ml:74-81) on the GPU, on a real MH chain (test/mcmc_test.ml:100-112)"""
    g = P.gauss_diag([0.4], [1.3])
    ctx.set_seed(8)
    rows = mcmc.mcmc_array(1000, g.scaled(0.75), g.scaled(0.25), P.box_proposal([1.3]), [0.4], ctx=ctx).chain(0)
    rows = np.ascontiguousarray(rows)
    out = np.empty_like(rows)
    k = C.c_int64()
    ctx.check(ctx.lib.mg_remove_repeat_samples(ctx.h, _abi.ptr(rows), C.c_int64(rows.shape[0]), C.c_int32(1), _abi.ptr(out), C.byref(k)))
    got = out[: k.value]
    assert np.array_equal(got, og.remove_repeat_samples(rows, 1))
    assert np.array_equal(got, mcmc.remove_repeat_samples(rows, 1))
    assert np.all(got[1:, 0] != got[:-1, 0]) and 1 < len(got) < len(rows)


def test_memory_helpers_and_stream(ctx):
    p = C.c_void_p()
    ctx.check(ctx.lib.mg_malloc_device(ctx.h, C.c_int64(1 << 20), C.byref(p)))
    h = C.c_void_p()
    ctx.check(ctx.lib.mg_malloc_pinned(ctx.h, C.c_int64(1 << 20), C.byref(h)))
    src = np.arange(1 << 17, dtype=np.float64)
    C.memmove(h, src.ctypes.data, src.nbytes)
    ctx.check(ctx.lib.mg_memcpy_h2d(ctx.h, p, h, C.c_int64(src.nbytes)))
    back = np.empty_like(src)
    ctx.check(ctx.lib.mg_memcpy_d2h(ctx.h, back.ctypes.data_as(C.c_void_p), p, C.c_int64(src.nbytes)))
    assert np.array_equal(back, src)
    ctx.check(ctx.lib.mg_free_pinned(ctx.h, h))
    ctx.check(ctx.lib.mg_free_device(ctx.h, p))
    assert ctx.launch_count >= 0 and ctx.last_kernel_ms >= 0.0
    tf, gbs = C.c_double(), C.c_double()
    ctx.check(ctx.lib.mg_measure_fp64_tflops(ctx.h, 1, C.byref(tf)))
    ctx.check(ctx.lib.mg_measure_store_gbs(ctx.h, C.c_int64(1 << 28), 1, C.byref(gbs)))
    assert tf.value > 1.0 and gbs.value > 100.0


GAUSS_BODY = """
  double s = 0.0;
  for (int i = 0; i < dim; ++i) {
    const double d = (x[i] - p[i]) / p[dim + i];
    s = s + (-0.91893853320467274178 - log(p[dim + i]) - 0.5 * d * d);
  }
  return s + 0.0;
"""


def test_runtime_plugin_matches_builtin(ctx, og):
    """mg_plugin_register_source: a user log-density compiled with NVRTC and
    inlined into the sampler kernel; written with the arithmetic of
    Stats.log_multi_gaussian it must reproduce the built-in plugin bit for bit"""
    from mcmc_ocaml_b200 import InvalidArgument
    D = 3
    mu, sig = [0.3, 0.5, 0.7], [0.1, 0.2, 0.05]
    user = P.register_source("my_gaussian", GAUSS_BODY, D, np.concatenate([mu, sig]), ctx=ctx)
    assert user.kind >= 1000
    builtin = P.gauss_diag(mu, sig)
    x = np.random.default_rng(0).random((500, D))
    assert np.array_equal(eval_gpu(ctx, user, x), eval_gpu(ctx, builtin, x))
    prior = P.box(np.zeros(D), np.ones(D), 0.0)
    prop = P.wrap_proposal(np.zeros(D), np.ones(D), [0.2, 0.3, 0.1])
    ctx.set_seed(5)
    a = mcmc.mcmc_array(80, user, prior, prop, [0.3, 0.5, 0.7], nchains=200, nbin=7, nskip=2, ctx=ctx)
    ctx.set_seed(5)
    b = mcmc.mcmc_array(80, builtin, prior, prop, [0.3, 0.5, 0.7], nchains=200, nbin=7, nskip=2, ctx=ctx)
    assert np.array_equal(a.block, b.block) and np.array_equal(a.accept, b.accept)
    # a second registration (module recompiled with both functions); the first keeps working
    banana = P.register_source("banana", "const double a = x[1] - x[0]*x[0]; return -0.5*(x[0]*x[0] + 4.0*a*a);", 2, ctx=ctx)
    r = mcmc.mcmc_array(200, banana, P.zero(2), P.box_proposal([0.5, 0.5]), [0.0, 0.0], nchains=1024, nbin=2000, nskip=20, ctx=ctx)
    xs = r.values()
    assert abs(xs[:, 0].mean()) < 0.05 and abs(xs[:, 0].std() - 1.0) < 0.08      # x0 ~ N(0,1) marginally
    assert abs((xs[:, 1] - xs[:, 0] ** 2).std() - 0.5) < 0.04                   # x1 - x0^2 ~ N(0, 1/4)
    assert np.array_equal(eval_gpu(ctx, user, x), eval_gpu(ctx, builtin, x))
    with pytest.raises(InvalidArgument):                                        # does not compile
        P.register_source("broken", "return undefined_symbol;", 2, ctx=ctx)


def test_runtime_plugins_in_rjmcmc_and_nested(ctx, og):
    """User log-densities (NVRTC) as the closures of Mcmc.rjmcmc_array (mcmc.mli:132-140) and Nested.nested_evidence
    (nested.mli:50-61): written with the built-in plugins' arithmetic they give the built-ins' runs value for value."""
    from mcmc_ocaml_b200 import interpolate_pdf, nested
    D = 3
    mu, sig = [0.3, 0.5, 0.7], [0.1, 0.2, 0.05]
    user = P.register_source("my_gaussian_again", GAUSS_BODY, D, np.concatenate([mu, sig]), ctx=ctx)
    builtin = P.gauss_diag(mu, sig)
    # ---- Nested: user likelihood under a built-in box prior, and the other way round
    box_body = ("for (int i = 0; i < dim; ++i) { if (!(x[i] > p[i] && x[i] < p[dim + i])) return -__longlong_as_double(0x7FF0000000000000ll); }"
                " return p[2 * dim];")
    ubox = P.register_source("my_open_box", box_body, D, np.concatenate([np.zeros(D), np.ones(D), [0.0]]), ctx=ctx)
    prior = P.box(np.zeros(D), np.ones(D), 0.0, closed=False)
    runs = []
    for like, pr in ((builtin, prior), (user, prior), (builtin, ubox), (user, ubox)):
        ctx.set_seed(404)
        runs.append(nested.nested_evidence(like, pr, np.zeros(D), np.ones(D), nlive=200, nmcmc=25, batch=32, ctx=ctx))
    for r in runs[1:]:
        assert np.array_equal(r.points, runs[0].points) and np.array_equal(r.log_likelihood, runs[0].log_likelihood)
        assert r.log_evidence == runs[0].log_evidence
    # ---- RJMCMC: user likelihood in model A
    rng = np.random.default_rng(3)
    pts = rng.normal(mu, sig, (20000, D)).clip(0.0, 1.0)
    ip = interpolate_pdf.InterpPdf(pts, np.zeros(D), np.ones(D), ctx=ctx)
    prop = P.wrap_proposal(np.zeros(D), np.ones(D), [0.05, 0.1, 0.03])
    pb = P.box(np.zeros(D), np.ones(D), -0.7)
    out = []
    for like in (builtin, user):
        A = mcmc.RjModel(like, P.box(np.zeros(D), np.ones(D), 0.0), prop, 0.5, interp=ip)
        B = mcmc.RjModel(builtin, pb, prop, 0.5, interp=ip)
        ctx.set_seed(505)
        out.append(mcmc.rjmcmc_array(60, A, B, mu, mu, nskip=2, nbin=5, nchains=300, record_samples=True, ctx=ctx))
    assert np.array_equal(out[0].model, out[1].model) and np.array_equal(out[0].block, out[1].block)
    assert out[0].counts == out[1].counts and out[0].cross == out[1].cross and out[0].cross[1] > 0


def test_radix_sort_sorted_and_stable(ctx):
    """the hand-written radix sort under Kd_tree / Evidence: sorted and stable at ragged sizes,
    with ties, signed zeros, negative values and constant-digit passes"""
    import torch
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    cases = [torch.rand(n, dtype=torch.float64, device="cuda", generator=g) for n in (2, 4095, 4096, 4097, 12293, 1 << 20, (1 << 20) + 1)]
    cases.append(torch.randn(300001, dtype=torch.float64, device="cuda", generator=g) * 1e6)
    cases.append(torch.randint(0, 7, (200000,), device="cuda", generator=g).double() - 3.0)      # heavy ties
    cases.append(torch.tensor([0.0, -0.0, 0.0, -0.0, 1.0, -1.0] * 5000, dtype=torch.float64, device="cuda"))
    cases.append(torch.full((70000,), 0.5, dtype=torch.float64, device="cuda"))                   # every pass skipped
    torch.cuda.synchronize()
    for v in cases:
        k = C.c_int64(-1)
        ctx.check(ctx.lib.mg_debug_sort_check(ctx.h, C.c_void_p(v.data_ptr()), C.c_int64(v.numel()), C.byref(k)))
        assert k.value == 0, v.numel()


def test_bulk_draws_match_oracle(ctx, og):
    """Stats.draw_uniform / draw_gaussian / draw_cauchy (stats.ml:89-91,113-128): same Philox stream on both sides."""
    from mcmc_ocaml_b200 import stats
    n = 200_000
    for kind, fn, a, b in [(0, stats.draw_uniform, -2.0, 3.5), (1, stats.draw_gaussian, 0.7, 1.3), (2, stats.draw_cauchy, -1.0, 0.25)]:
        ctx.set_seed(4242)
        e0 = ctx.epoch
        got = fn(a, b, n, ctx=ctx)
        want = og.stats_draw(4242, e0, kind, a, b, n)
        if kind == 2:   # tan: CUDA libm vs glibc
            np.testing.assert_allclose(got, want, rtol=1e-9)
        else:
            assert np.array_equal(got, want)
    g = stats.draw_gaussian(0.7, 1.3, n, ctx=ctx)
    assert abs(g.mean() - 0.7) < 0.02 and abs(g.std() - 1.3) < 0.02       # test/stats_test.ml style moments
    u = stats.draw_uniform(-2.0, 3.5, n, ctx=ctx)
    assert u.min() >= -2.0 and u.max() < 3.5 and abs(u.mean() - 0.75) < 0.02


def test_accept_prefilter_decides_like_float64(ctx):
    """rng.cuh log_u_less_than: the single-precision estimate may only decide when it cannot be wrong.  Thresholds
    are placed at log u +- k ulps, +- 1e-7 .. 1e-3 relative (inside and outside the tolerance band), and at the
    special values; the fast decision must equal the float64 comparison in every case."""
    rng = np.random.default_rng(5)
    base = np.concatenate([rng.random(200_000), rng.random(50_000) * 1e-3, 1.0 - rng.random(50_000) * 1e-9,
                           np.array([0.0, 5e-324, 1e-310, 1e-40, 1e-31, 1e-30, 1e-29, 2.0 ** -52, 0.5, 1.0 - 2.0 ** -53])])
    with np.errstate(divide="ignore"):
        lg = np.log(base)
    us, ds = [], []
    for rel in [0.0, 1e-16, 1e-12, 1e-9, 1e-7, 3e-6, 1e-5, 3e-5, 1e-4, 1e-3, 0.1]:
        for sign in (-1.0, 1.0):
            us.append(base); ds.append(lg + sign * rel * (1.0 + np.abs(lg)))
    for k in (-2, -1, 1, 2):
        us.append(base); ds.append(np.nextafter(lg, np.inf * k) if abs(k) == 1 else np.nextafter(np.nextafter(lg, np.inf * k), np.inf * k))
    for special in (np.inf, -np.inf, np.nan, 0.0, -1e308, 1e308, -745.0, -800.0):
        us.append(base); ds.append(np.full(base.size, special))
    u = np.ascontiguousarray(np.concatenate(us)); d = np.ascontiguousarray(np.concatenate(ds))
    fast = np.empty(u.size, np.uint8); exact = np.empty(u.size, np.uint8)
    ctx.check(ctx.lib.mg_debug_accept_test(ctx.h, _abi.ptr(u), _abi.ptr(d), C.c_int64(u.size),
                                           fast.ctypes.data_as(C.c_void_p), exact.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(fast, exact)
    assert 0 < exact.sum() < exact.size


def test_trim_pool_releases_cached_memory(ctx):
    """mg_ctx_trim_pool: temporaries cached in the stream-ordered pool go back to the driver; results unchanged after"""
    import torch
    from mcmc_ocaml_b200 import evidence
    ll = np.random.default_rng(0).normal(-3, 1, 2_000_000)
    z0 = evidence.evidence_harmonic_mean(ll=ll, ctx=ctx)
    free_cached, _ = torch.cuda.mem_get_info()
    ctx.trim_pool()
    free_trimmed, _ = torch.cuda.mem_get_info()
    assert free_trimmed >= free_cached
    assert evidence.evidence_harmonic_mean(ll=ll, ctx=ctx) == z0


def test_bounded_wait_reports_error_and_context_survives(ctx, og, monkeypatch):
    """The balanced sampler's queue waits are bounded.  Running out of the budget must not trap (a trap poisons the
    CUDA context of the whole process): the call returns MG_ECUDA (Failure "cuda: ...") and the SAME context then
    runs the same ensemble correctly.  The time-out is provoked by a persistent grid four times larger than the
    number of chain groups (the surplus warps must wait for the first segments to finish) and a budget of 1,000
    cycles."""
    D = 4
    mu = np.arange(D) / 10.0
    cov = 0.7 ** np.abs(np.subtract.outer(np.arange(D), np.arange(D)))
    like, prior, prop = P.gauss_corr(mu, cov), P.zero(D), P.box_proposal(np.full(D, 0.5))
    C_, n = 592 * 32 + 64, 600
    monkeypatch.setenv("MCMC_GPU_MH_GRID", str(4 * (C_ // 32)))
    monkeypatch.setenv("MCMC_GPU_WAIT_CYCLES", "1000")
    ctx.set_seed(5)
    with pytest.raises(Failure, match="timed out"):
        mcmc.mcmc_array(n, like, prior, prop, mu, nchains=C_, ctx=ctx)
    monkeypatch.delenv("MCMC_GPU_MH_GRID")
    monkeypatch.delenv("MCMC_GPU_WAIT_CYCLES")
    ctx.set_seed(5)
    got = mcmc.mcmc_array(n, like, prior, prop, mu, nchains=C_, ctx=ctx)          # same context, healthy
    want, acc, _ = og.mcmc_array(5, 0, n, like, prior, prop, mu, nchains=C_, nthreads=16)
    assert np.array_equal(got.block, want) and np.array_equal(got.accept, acc)


def test_blob_validation_rejects_corrupt_headers(ctx):
    """mg_kdtree_from_blob_dev trusts nothing in a blob that arrived from another process"""
    import torch

    from mcmc_ocaml_b200 import InvalidArgument, kd_tree
    pts = np.random.default_rng(1).random((500, 3))
    t = kd_tree.KdTree(pts, np.zeros(3), np.ones(3), ctx=ctx)
    p, nbytes = t.blob()
    buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    ctx.check(ctx.lib.mg_memcpy_d2d(ctx.h, C.c_void_p(buf.data_ptr()), C.c_void_p(p), C.c_int64(nbytes)))
    good = buf.clone()
    hdr = good[:128].cpu().numpy().copy()
    i64 = hdr.view(np.int64)
    # header layout (csrc/kdtree.cuh KdHeader): magic, N, nnodes, nbytes, (D, nlevels), (min_split, pad), off_*[7]
    for field, value in ((1, 1 << 40), (2, 10 * 500), (6, 64), (10, nbytes + 256), (11, 8)):
        bad = i64.copy(); bad[field] = value
        b2 = good.clone(); b2[:128] = torch.as_tensor(bad.view(np.uint8), device="cuda")
        torch.cuda.synchronize()
        with pytest.raises(InvalidArgument):
            kd_tree.KdTree.from_blob(b2.data_ptr(), nbytes, ctx=ctx)
    torch.cuda.synchronize()
    t2 = kd_tree.KdTree.from_blob(good.data_ptr(), nbytes, ctx=ctx)          # the untouched copy still loads
    assert t2.nnodes == t.nnodes


def test_c_program_drives_the_abi():
    """tests/c/test_capi.c: the call sequences of ocaml/mcmc_gpu_stubs.c (mg_mcmc_array, Interp, mg_rjmcmc_array,
    Evidence, Stats, mg_nested_evidence, mg_rjmcmc_array_k, mg_ellipse_*, the pool calls) from plain C, no Python between
    the program and libmcmcgpu.so; checks the reference's known answers (ratio 4.0 +- 0.1, nested evidence 1 within 2x
    its error, every point inside its enclosing ellipse) and that two models through the k-model call equal the
    two-model call."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(here, "c", "test_capi")
    if not os.path.exists(exe):
        subprocess.check_call(["gcc", "-O1", "-I" + os.path.join(here, "..", "include"), os.path.join(here, "c", "test_capi.c"),
                               "-o", exe, "-L" + os.path.join(here, "..", "mcmc_ocaml_b200"), "-lmcmcgpu", "-lm",
                               "-Wl,-rpath,$ORIGIN/../../mcmc_ocaml_b200"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok:")


def test_quadform_on_fp64_tensor_cores(ctx):
    """The decision experiment of DESIGN.md (FP64 tensor cores for the high-D Gaussian quadratic form): the DMMA
    variant agrees with the FMA variant to 1e-13 relative (a 4-term dot product inside mma.sync.m8n8k4.f64 is summed
    in an unspecified order), the FMA variant IS the sampler plugin's arithmetic (identical to MG_FN_GAUSS_CORR)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import dmma_decision
    doc = dmma_decision.run(M=100_003, reps=1)
    for c in doc["cases"]:
        assert c["fma_identical_to_plugin"] and c["max_rel_diff_dmma_vs_fma"] < 1e-13
    assert doc["fp64_dmma_tflops"] > 1.0


def test_pool_trim_and_reserve(ctx):
    """mg_ctx_trim_pool hands the cached temporaries back, mg_ctx_reserve_pool makes the pool hold one large piece again;
    results do not depend on either."""
    from mcmc_ocaml_b200 import evidence
    rng = np.random.default_rng(3)
    ll = rng.normal(-3.0, 1.0, 100_000)
    z0 = evidence.evidence_harmonic_mean(ll=ll, ctx=ctx)
    ctx.trim_pool()
    assert evidence.evidence_harmonic_mean(ll=ll, ctx=ctx) == z0
    ctx.reserve_pool(1.0)
    ctx.reserve_pool(0.0)                      # nothing to do
    assert evidence.evidence_harmonic_mean(ll=ll, ctx=ctx) == z0
    ctx.reserve_pool()                         # back to the default for the tests that follow
