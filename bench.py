#!/usr/bin/env python
"""bench.py -- sharp-null simulations/sec at n = 4 000 units/side (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload = "C4"): NYC-2015-shaped synthetic district pair, 4 000 units/side (n = 8 000),
Matern-3/2 + Const kernel, 100 border sentinels, hypotest_mll null with 20 000 draws per GPU.
One STEP = one pass of the hot path over one batch of 20 000 draws on every rank: Philox normals ->
y = L_N z (prior_rand) -> alpha_T, alpha_C, alpha_N (update_alpha!) -> the three mll's -> exceedance tally
(src/hypotest_mll.jl:1-45), all through the C ABI (geordd_null_mll).  Factors are resident in HBM when the
timed region starts (`value`); `e2e` repeats the measurement through the public host API from HOST buffers
(fit T, C and the pooled null from pinned host X / y, 20 000 draws, statistics copied back).

Multi-GPU: one process per GPU, draws sharded by index (weak scaling: 20 000 draws per GPU per step), one
int64 all-reduce of the tally per step over NCCL; time = max over ranks of the device time.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_SIDE, NB, S_DRAWS = 4000, 100, 20000
ELL, SIG2, CONST_S2, NOISE_SD = 0.3, 1.0, 400.0, 0.3
METRIC, UNIT = "sharp_null_simulations_per_sec", "sims/s"
DMMA_PEAK_UNTHROTTLED = 37.07   # TFLOP/s, mma.sync.m8n8k4.f64 register-resident loop on an unthrottled B200 of this pool


def flops_per_mll_draw(m):      # n^2 (TRMV) + 2 m_T^2 + 2 m_C^2 + 2 n^2 = 16 m^2
    return 16.0 * m * m


def synth(O=None):
    """SURVEY §8d synthetic two-region data (shape-faithful to test/runtests.jl:4-36), seed 1."""
    rng = np.random.default_rng(1)
    n = 2 * M_SIDE
    X = rng.random((2, n))
    rad = np.sqrt(X[0] ** 2 + X[1] ** 2)
    order = np.argsort(rad, kind="stable")
    X, rad = X[:, order], rad[order]
    r = 0.5 * (rad[M_SIDE - 1] + rad[M_SIDE])
    treat = np.arange(n) < M_SIDE
    Y = (-0.8 + np.sin(X[0] * 2.0) + 3.0 * X[1] ** 2 + 0.3 * rng.standard_normal(n) + 0.3 * treat)
    th = np.linspace(0.0, np.pi / 2.0, NB)
    Xb = np.asfortranarray(np.vstack([r * np.cos(th), r * np.sin(th)]))
    return (np.asfortranarray(X[:, :M_SIDE]), np.asfortranarray(X[:, M_SIDE:]), Y[:M_SIDE].copy(), Y[M_SIDE:].copy(), Xb)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(sample_draws, threads=None):
    """The reference's per-draw CPU routine sequence (dtrmv, dpotrs x3, dots -- src/hypotest_mll.jl:1-19) on the
    oracle, one draw at a time exactly as the Julia package does; bounded sample of the C4 workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import geordd_oracle as O
    XT, XC, YT, YC, Xb = synth()
    spec = O.KernelSpec(O.MAT32, ELL, SIG2, CONST_S2, 0.0)
    ln = math.log(NOISE_SD)
    t0 = time.time()
    gT, gC = O.gp_fit(spec, XT, YT, ln), O.gp_fit(spec, XC, YC, ln)
    gN = O.make_null(gT, gC)
    setup_s = time.time() - t0
    Z = np.random.default_rng(2).standard_normal((2 * M_SIDE, sample_draws + 1))
    O.nsim_logP(gT, gC, gN, Z[:, :1])                      # warm-up draw
    t0 = time.time()
    O.nsim_logP(gT, gC, gN, Z[:, 1:])
    dt = time.time() - t0
    return sample_draws / dt, setup_s, dt


def run_reference(args, rank, world):
    """--impl reference: the reference CPU path (Julia is absent -> the oracle port of it) on this box's host cores.
    Each step is a bounded sample (24 draws) of the C4 workload; rank 0 alone runs it."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import geordd_oracle as O
    cores = os.cpu_count()
    sample = 24
    XT, XC, YT, YC, Xb = synth()
    spec = O.KernelSpec(O.MAT32, ELL, SIG2, CONST_S2, 0.0)
    ln = math.log(NOISE_SD)
    t0 = time.time()
    gT, gC = O.gp_fit(spec, XT, YT, ln), O.gp_fit(spec, XC, YC, ln)
    gN = O.make_null(gT, gC)
    setup_s = time.time() - t0
    rng = np.random.default_rng(3)
    vals = []
    for i in range(args.warmup + args.steps):
        Z = rng.standard_normal((2 * M_SIDE, sample if i >= args.warmup else 2))
        t0 = time.time()
        O.nsim_logP(gT, gC, gN, Z)
        if i >= args.warmup:
            vals.append(sample / (time.time() - t0))
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: NYC-2015-shaped synthetic district pair, 4000 units/side, Matern-3/2 + Const, "
                                   "hypotest_mll null draws", "m_per_side": M_SIDE, "draws_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} mll null draws per step at n=8000 (oracle port of src/hypotest_mll.jl:1-19 via "
                                       f"SciPy/OpenBLAS dtrmv + 3 dpotrs per draw, all {cores} host threads); setup (assemble + "
                                       f"3 dpotrf) {setup_s:.1f} s excluded"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="geordd_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--draws", type=int, default=S_DRAWS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import georddb200
    G = georddb200.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = G.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    S = args.draws

    # ---- setup (untimed): fits and draw-independent state, resident in HBM --------------------------------
    XT, XC, YT, YC, Xb = synth()
    k = G.Mat32Iso(math.log(ELL), 0.5 * math.log(SIG2)) + G.Const(0.5 * math.log(CONST_S2))
    ln = math.log(NOISE_SD)
    gT, gC = G.GPE(XT, YT, G.MeanZero(), k, ln, ctx=ctx), G.GPE(XC, YC, G.MeanZero(), k, ln, ctx=ctx)
    gN = G.NullGPE(gT, gC, Xb, 0.0)
    d_obs = gT.mll + gC.mll - gN.mll
    import ctypes as C
    lib = ctx.lib
    tally = torch.zeros(1, dtype=torch.int64, device=dev)

    def step(i):
        cnt = C.c_longlong()
        draw0 = (i * world + rank) * S          # disjoint draw indices per (step, rank)
        ctx.check(lib.geordd_null_mll(gN._h, S, 1, draw0, None, d_obs, None, None, C.byref(cnt), None, None))
        if world > 1:
            tally[0] = cnt.value
            dist.all_reduce(tally)              # the single tiny NCCL all-reduce of the exceedance tally
        return cnt.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local_rank)
    launches0 = lib.geordd_launch_count()
    ctx.profile(True)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = lib.geordd_launch_count() - launches0
    prof = {nm: ctx.profile_get(nm) for nm in ("solve_seq", "solve_fwd", "solve_bwd", "trmm", "rng", "finish", "gemm", "potrf", "cov")}
    ctx.profile(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * S * args.steps / (ms * 1e-3)

    # ---- end to end through the public host API from pinned HOST buffers ----------------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = a
        return np.asfortranarray(v) if v.ndim == 1 else v, t

    # (2 x m) arrays are passed column-major; keep them as pinned 1-d storage reshaped in Fortran order
    def pinned_f(a):
        t = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        return v, t

    hXT, _k1 = pinned_f(XT); hXC, _k2 = pinned_f(XC); hYT, _k3 = pinned(YT); hYC, _k4 = pinned(YC)
    out_null = torch.empty(S, dtype=torch.float64, pin_memory=True)
    out_alt = torch.empty(S, dtype=torch.float64, pin_memory=True)

    def e2e_step(i):
        a = G.GPE(hXT, hYT, G.MeanZero(), k, ln, ctx=ctx)          # H2D coords + y, assemble, factor, alpha, mll
        b = G.GPE(hXC, hYC, G.MeanZero(), k, ln, ctx=ctx)
        nn = G.NullGPE(a, b)                                        # pooled null fit (n = 8000)
        dd = a.mll + b.mll - nn.mll
        cnt = C.c_longlong()
        draw0 = ((1000 + i) * world + rank) * S
        ctx.check(lib.geordd_null_mll(nn._h, S, 1, draw0, None, dd, G.capi.dptr(out_null.numpy()), G.capi.dptr(out_alt.numpy()),
                                      C.byref(cnt), None, None))
        if world > 1:
            tally[0] = cnt.value
            dist.all_reduce(tally)
        p = cnt.value / S
        nn.close(); a.close(); b.close()
        return p

    e2e_step(0)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(2, min(args.steps, 3))
    t_wall0 = time.time()
    f0.record(stream)
    for i in range(n_e2e):
        e2e_step(1 + i)
    f1.record(stream)
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), (time.time() - t_wall0) * 1e3)   # host work between calls counts too
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_val = world * S * n_e2e / (ms_e2e * 1e-3)
    h2d = 8 * (XT.size + XC.size + YT.size + YC.size) * 2            # T, C and the pooled null each upload X and y
    d2h = 8 * (2 * S + 2 * M_SIDE * 2) + 8 * 8                       # statistics + alphas + tallies

    if rank == 0:
        # ---- extra measurements on rank 0: FP64 tensor-pipe ceiling, Cholesky, cliff-face latency -------------
        dmma_live = ctx.dmma_peak_tflops()
        # The register-resident DMMA loop draws more power than any real kernel; on a power-capped box it measures
        # below what the solve kernels themselves sustain.  The roofline denominator is therefore the larger of the
        # live probe and the unthrottled probe value measured on this pool in round 1 (37.07 TFLOP/s =
        # 148 SMs x 128 FP64 tensor flop/clk x 1.965 GHz to 0.4 %; profiles/r01_bench_first.json).
        dmma_peak = max(dmma_live, DMMA_PEAK_UNTHROTTLED)
        potrf_ms, _ = ctx.bench_potrf(8192, 2)
        t0 = time.time()
        for _ in range(3):
            a = G.GPE(hXT, hYT, G.MeanZero(), k, ln, ctx=ctx); b = G.GPE(hXC, hYC, G.MeanZero(), k, ln, ctx=ctx)
            G.cliff_face(a, b, Xb)
            a.close(); b.close()
        cliff_fit_ms = (time.time() - t0) / 3 * 1e3
        t0 = time.time()
        for _ in range(3):
            G.cliff_face(gT, gC, Xb)
        cliff_only_ms = (time.time() - t0) / 3 * 1e3
        cnt = C.c_longlong()
        for rep in range(2):   # first call allocates the sentinel-system buffers of the handle
            ctx.sync(); t0 = time.time()
            ctx.check(lib.geordd_null_chi2(gN._h, S, 1, 10 ** 9, None, 1e300, None, C.byref(cnt)))
            ctx.sync(); chi_rate = S / (time.time() - t0)
        # opt-in forward-only evaluation of the same mll statistic (|L^-1 (y-m)|^2 and z'z; not the headline path)
        ctx.set_mll_forward_only(True)
        try:
            for rep in range(2):
                ctx.sync(); t0 = time.time()
                ctx.check(lib.geordd_null_mll(gN._h, S, 1, 10 ** 9, None, d_obs, None, None, C.byref(cnt), None, None))
                ctx.sync(); fwd_only_rate = S / (time.time() - t0)
        finally:
            ctx.set_mll_forward_only(False)

        # Dominant kernel: the fused forward/backward solve sequence = 12 m^2 of the 16 m^2 flop of a draw (2 m^2 per
        # triangular-solve pair against L_T and L_C, 8 m^2 against L_N); the triangular multiply (4 m^2) is reported beside it.
        seq_ms, seq_n = prof["solve_seq"]
        trmm_ms, trmm_n = prof["trmm"]
        dmma_ms = seq_ms + trmm_ms + prof["solve_fwd"][0] + prof["solve_bwd"][0]
        seq_flops = 12.0 * M_SIDE * M_SIDE * S            # per launch (one launch per step)
        trmm_flops = 4.0 * M_SIDE * M_SIDE * S
        achieved = seq_flops * seq_n / (seq_ms * 1e-3) / 1e12
        traffic = None
        try:   # per-launch DRAM traffic of that kernel from the committed ncu capture (same draw count only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01d_traffic.json")))
            if tj.get("draws") == S:
                traffic = tj["dram_gbytes"]["solve_seq_kernel"] * 1e9
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": "solve_seq_kernel (forward + backward solves against L_T, L_C, L_N of one batch of draws, FP64 DMMA)",
                    "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
                    "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                    "launches": int(seq_n), "avg_launch_ms": seq_ms / max(seq_n, 1),
                    "algorithmic_flops_per_launch": seq_flops,
                    "peak_source": "max(live register-resident mma.sync.m8n8k4.f64 probe, 37.07 = the same probe unthrottled on this pool); "
                                   "MEASURED_PEAKS.json has no FP64 figure", "peak_live": dmma_live,
                    "share_of_step": seq_ms / ms,
                    "also": {"kernel": "solve_dataflow_kernel<TRMM> (prior_rand, 4 m^2 flop/draw)",
                             "achieved": trmm_flops * trmm_n / (trmm_ms * 1e-3) / 1e12, "avg_launch_ms": trmm_ms / max(trmm_n, 1),
                             "share_of_step": trmm_ms / ms},
                    "all_dmma_launches": {"achieved": flops_per_mll_draw(M_SIDE) * S * args.steps / (dmma_ms * 1e-3) / 1e12,
                                          "share_of_step": dmma_ms / ms}}
        cpu = None
        if not args.no_cpu_baseline:
            rate, setup_s, dt = cpu_reference_rate(16)
            cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"16 mll null draws at n=8000 ({dt:.1f} s) after {setup_s:.1f} s setup; oracle port of "
                             f"src/hypotest_mll.jl:1-19 (dtrmv + 3 dpotrs per draw, SciPy/OpenBLAS, all host threads)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C4: NYC-2015-shaped synthetic district pair, 4000 units/side, Matern-3/2 + Const, "
                                       "hypotest_mll null, 20000 draws per GPU per step",
                           "m_per_side": M_SIDE, "n": 2 * M_SIDE, "sentinels": NB, "draws_per_gpu_per_step": S,
                           "rng": "device Philox-4x32-10 + Box-Muller", "parallelism": f"draw-sharded x{world}",
                           "l2": "inputs larger than L2 (factors 768 MB + 6.6 GB of draw buffers per step)"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "what": "per step: GPE(T), GPE(C), make_null from pinned host X/y + 20000 mll draws + D2H of both "
                                "statistic vectors (boot_mlltest through the host API)", "steps": n_e2e},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "cpu_baseline": cpu,
                "extra": {"chi2_sims_per_sec_1gpu": chi_rate, "mll_forward_only_sims_per_sec_1gpu": fwd_only_rate, "cliff_face_fit_ms_4k_per_side": cliff_fit_ms,
                          "cliff_face_only_ms_4k_per_side": cliff_only_ms, "potrf_8192_ms": potrf_ms,
                          "potrf_8192_tflops": 8192 ** 3 / 3 / (potrf_ms * 1e-3) / 1e12, "dmma_peak_tflops": dmma_peak,
                          "profile_ms": {k_: v_[0] for k_, v_ in prof.items()}}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
