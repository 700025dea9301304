#!/usr/bin/env python
"""bench.py -- headline benchmark of the sampling path (BASELINE.json config 2).

Workload ("step" = one pass of the hot path over one batch): 65,536
independent Metropolis-Hastings chains on a 10-D correlated Gaussian
(mu_i = i/10, Sigma_ij = 0.7^|i-j|), box proposal h = 0.5, flat prior,
10,000 steps per chain, every step recorded (nskip = 1) into a
[n][D+2][C] float64 sample block in HBM (62.9 GB per pass per GPU).
Metric: chain-steps/s, whole job over all GPUs (weak scaling: every rank
runs its own 65,536 chains with global chain ids rank*C ...).

  value  device-resident: state and sample block live in HBM, one kernel.
  e2e    the C-ABI call mg_mcmc_array_resident with HOST buffers: start
         points copied from pinned host memory, the sample block kept in HBM
         for the GPU consumers that follow, and final states + accept counts
         + per-field mean/std over all 6.6e8 recorded samples read back.
  --impl reference   the CPU restatement of the OCaml reference
         (oracle/, the OCaml toolchain does not exist in this image) on all
         host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 10
NCHAINS = 65536
NSTEPS = 10000          # steps per chain per pass; n = NSTEPS + 1 samples, nskip = 1
SEED = 0x5EED0001
BYTES_PER_STEP = 8 * (D + 2)   # one recorded sample per chain-step at nskip = 1


def model():
    from mcmc_ocaml_b200 import plugins as P
    mu = np.arange(D) / 10.0
    cov = 0.7 ** np.abs(np.subtract.outer(np.arange(D), np.arange(D)))
    return mu, P.gauss_corr(mu, cov), P.zero(D), P.box_proposal(np.full(D, 0.5))


def workload_config(Cn=NCHAINS, T=NSTEPS):
    n, F = T + 1, D + 2
    return {"workload": f"cfg2: {Cn} independent MH chains per GPU, {D}-D correlated Gaussian "
                        f"(Sigma_ij=0.7^|i-j|), box proposal h=0.5, {T} steps each, nskip=1, "
                        f"every sample recorded ([n][D+2][C] f64, {n * F * Cn * 8 / 1e9:.1f} GB/pass)",
            "chains_per_gpu": Cn, "dim": D, "steps_per_chain": T, "nskip": 1,
            "l2": "outputs (62.9 GB/pass) far exceed the 126 MB L2; no flush needed",
            "rng": "Philox4x32-10, 52-bit uniforms, 6 blocks/step"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag.set()
        if self.t:
            self.t.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(nthreads, target_seconds=12.0):
    """Time the CPU restatement on a bounded sample of the workload."""
    from oracle import oracle as og
    mu, like, prior, prop = model()
    og.lib()
    # calibrate: one short run, then size the sample for ~target_seconds
    c0, n0 = max(nthreads * 4, 8), 201
    t = time.perf_counter()
    og.mcmc_array(SEED, 0, n0, like, prior, prop, mu, nchains=c0, nthreads=nthreads, record=False)
    dt = time.perf_counter() - t
    rate = c0 * (n0 - 1) / dt
    # the sample keeps the real chain length (1e4 steps) and bounds the number of chains
    steps = NSTEPS
    chains = int(max(nthreads, min(NCHAINS, rate * target_seconds / steps)))
    chains = max(nthreads, (chains // nthreads) * nthreads)
    t = time.perf_counter()
    og.mcmc_array(SEED, 1, steps + 1, like, prior, prop, mu, nchains=chains, nthreads=nthreads, record=True)
    dt = time.perf_counter() - t
    return chains * steps / dt, f"{chains} chains x {steps} steps of the same model, samples recorded, {nthreads} threads"


def run_reference(args, rank, world):
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    vals, sample = [], ""
    per_step = max(2.0, min(12.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference(nthreads, per_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sample = cpu_reference(nthreads, per_step)
        vals.append(v)
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    print(json.dumps({
        "impl": "reference", "metric": "chain-steps/sec", "value": v, "unit": "chain-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args.chains, args.chain_steps),
                       note="CPU arm: each step is a bounded sample of this workload, see cpu_baseline.sample"),
        "cpu_baseline": {"value": v, "unit": "chain-steps/s", "cores": nthreads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of farr/mcmc-ocaml (oracle/), not OCaml: no OCaml toolchain here"},
        "e2e": {"value": v, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=NCHAINS)
    ap.add_argument("--chain-steps", type=int, default=NSTEPS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-evidence", action="store_true", help="skip the config-3 evidence leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from mcmc_ocaml_b200 import Context, _abi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    mu, like, prior, prop = model()
    Cn, T = args.chains, args.chain_steps
    n, F = T + 1, D + 2
    ctx = Context(local_rank, SEED)
    stream = torch.cuda.Stream(dev)      # a real (non-NULL) stream shared by torch and the library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)   # time with torch events on the launching stream
    lib = ctx.lib

    # device-resident buffers (sized for 180 GB HBM: 62.9 GB sample block)
    state = torch.empty((F, Cn), dtype=torch.float64, device=dev)
    state[:D] = torch.tensor(mu, dtype=torch.float64, device=dev)[:, None]
    samples = torch.empty((n, F, Cn), dtype=torch.float64, device=dev)
    accept = torch.zeros(Cn, dtype=torch.int32, device=dev)
    cfg = _abi.mg_mcmc_cfg(Cn, D, 0, 0, 1, n, rank * Cn, 0, 0)
    ls, ps, js = like.spec(), prior.spec(), prop.spec()

    def step_dev():
        ctx.check(lib.mg_mcmc_array_dev(ctx.h, C.byref(ls), C.byref(ps), C.byref(js), C.byref(cfg),
                                        C.c_void_p(state.data_ptr()), C.c_void_p(samples.data_ptr()),
                                        C.c_void_p(accept.data_ptr())))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_dev()
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    e0.record(stream)
    marks[0].record(stream)
    for i in range(args.steps):
        step_dev()
        marks[i + 1].record(stream)
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - l0
    # average launch duration of the dominant kernel: CUDA events on the
    # launching stream around every launch of the timed region
    k_ms = float(np.mean([marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]))
    k_ms_lib = ctx.last_kernel_ms     # the library's own events around the last launch
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = Cn * T * world / (ms_per_step * 1e-3)

    # ---- e2e: C-ABI call with host buffers -------------------------------
    x0 = torch.empty((Cn, D), dtype=torch.float64).pin_memory()
    x0[:] = torch.tensor(mu, dtype=torch.float64)
    final = torch.empty((Cn, F), dtype=torch.float64).pin_memory()
    acc_h = torch.empty(Cn, dtype=torch.int64).pin_memory()
    rej_h = torch.empty(Cn, dtype=torch.int64).pin_memory()
    mean_h = np.empty(F); std_h = np.empty(F)
    cfg_e = _abi.mg_mcmc_cfg(Cn, D, 0, 0, 1, n, rank * Cn, 0, 0)

    def step_e2e():
        ctx.check(lib.mg_mcmc_array_resident(
            ctx.h, C.byref(ls), C.byref(ps), C.byref(js), C.byref(cfg_e), C.c_void_p(x0.data_ptr()),
            C.c_void_p(samples.data_ptr()), C.c_void_p(final.data_ptr()), C.c_void_p(acc_h.data_ptr()),
            C.c_void_p(rej_h.data_ptr()), _abi.ptr(mean_h), _abi.ptr(std_h)))

    e2e_warm = min(args.warmup, 2)
    for _ in range(e2e_warm):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()      # returns after its D2H copies completed
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    e2e_kernel_ms = ctx.last_kernel_ms   # the sampler launch inside the last call (library events)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = Cn * T * world / (e2e_ms / args.steps * 1e-3)
    h2d = Cn * D * 8
    d2h = Cn * F * 8 + Cn * 4 + 2 * F * 8

    # gather a physics check from every rank (NCCL all_gather of a few bytes)
    acc_rate = torch.tensor([float(acc_h.sum()) / float(acc_h.sum() + rej_h.sum())], dtype=torch.float64, device=dev)
    rates = [torch.zeros_like(acc_rate) for _ in range(world)]
    if world > 1:
        dist.all_gather(rates, acc_rate)
    else:
        rates = [acc_rate]

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = BYTES_PER_STEP * Cn * T / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "mh_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        fp64 = C.c_double(0.0)
        store = C.c_double(0.0)
        lib.mg_measure_fp64_tflops(ctx.h, 3, C.byref(fp64))
        lib.mg_measure_store_gbs(ctx.h, C.c_int64(8 << 30), 3, C.byref(store))
        out = {
            "metric": "chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(Cn, T),
            "e2e": {"value": e2e_value, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps, "sampler_kernel_ms_in_call": e2e_kernel_ms,
                    "call": "mg_mcmc_array_resident: pinned x0 -> device, MH kernel (per-chain running moments kept in "
                            "registers), sample block stays in HBM, Stats mean/std pooled from the chain moments, "
                            "final states + counters + stats -> host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": f"{peak_kind} hbm_gbs (burst copy)",
                         "kernel": "mh_balanced_kernel<GaussCorr<10>,ZeroFn,BoxProp<10>,10>",
                         "kernel_ms": k_ms, "kernel_ms_last_launch_lib_events": k_ms_lib, "bytes_per_chain_step": BYTES_PER_STEP,
                         "store_only_peak_gbs": store.value, "fp64_fma_tflops_measured": fp64.value,
                         "fp64_flops_per_chain_step": D * D + 8 * D + 3,
                         "fp64_tflops_achieved": (D * D + 8 * D + 3) * Cn * T / (k_ms * 1e-3) / 1e12},
            "accept_rate_per_rank": [float(r.item()) for r in rates],
            "posterior_mean_err_max": float(np.max(np.abs(mean_h[:D] - mu))),
        }
        if world == 1 and not args.no_evidence:
            # second metric of BASELINE.json: evidence samples/s (config 3 on this GPU)
            try:
                del samples
                torch.cuda.empty_cache()
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_evidence
                ev = bench_evidence.run(bench_evidence._Args(reps=3, device=local_rank), ctx=ctx)
                out["evidence"] = {
                    "workload": "cfg3: Weinberg (Lebesgue) kd-tree evidence + harmonic mean, 1e7 synthetic 20-D "
                                "posterior samples, device resident, one GPU",
                    "lebesgue_samples_per_s": ev["lebesgue_samples_per_s"], "lebesgue_s": ev["lebesgue_s"],
                    "harmonic_s": ev["harmonic_s"], "direct_s": ev["direct_s"],
                    "kdtree_full_build_s": ev.get("tree_full_s"), "kdtree_nodes": ev.get("tree_full_nodes"),
                    "interp_jump_prob_per_s": ev.get("jump_prob_queries_per_s"), "interp_draw_per_s": ev.get("draw_per_s")}
            except Exception as e:  # the headline line must still be printed
                out["evidence"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu:
            nthreads = os.cpu_count() or 1
            v, sample = cpu_reference(nthreads, 12.0)
            out["cpu_baseline"] = {"value": v, "unit": "chain-steps/s", "cores": nthreads, "kind": "port",
                                   "sample": sample,
                                   "note": "C++ restatement of farr/mcmc-ocaml (oracle/), not OCaml"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
