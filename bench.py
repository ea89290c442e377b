#!/usr/bin/env python
"""bench.py -- sharp-null simulations/sec at n = 4 000 units/side (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (config.workload = "C4"): NYC-2015-shaped synthetic district pair, 4 000 units/side (n = 8 000),
Matern-3/2 + Const kernel, 100 border sentinels, hypotest_mll null with 20 000 draws per GPU.
One STEP = one pass of the hot path over one batch of 20 000 draws on every rank: Philox normals (generated inside the
triangular multiply) -> y = L_N z (prior_rand) -> alpha_T, alpha_C, alpha_N (update_alpha!) -> the three mll's ->
exceedance tally (src/hypotest_mll.jl:1-45), all through the C ABI (geordd_null_mll).  Factors are resident in HBM when
the timed region starts (`value`); `e2e` repeats the measurement through the public host API from HOST buffers
(GPRealisations fit of T and C and the pooled null fit from pinned host X / y, 20 000 draws, statistics copied back) and
reports the MEDIAN over >= 10 steps.

Multi-GPU: one process per GPU, draws sharded by index (weak scaling: 20 000 draws per GPU per step), one int64
all-reduce of the tally per step over NCCL; time = max over ranks of the device time.

`extra` carries the other shapes BASELINE.json names, each with its flop count and fraction of the FP64 tensor peak:
  c4_strong  boot_mlltest(gpT, gpC, 20 000) -- 20 000 draws IN TOTAL, sharded over the ranks, wall time per call from host
             buffers (pooled fit included): the honest strong-scaling number;
  c3         Mississippi-shaped sweep: 1 000 units/side, 200 sentinels, boot_chi2test + boot_invvar with 10 000 draws, sharded;
  c5         power study: 16 000 units, 1 000 null draws x 3 statistics + nsim_power 500 draws, sharded;
  c2         SimulatedExample-shaped joint fit: 2 000 units/side, p = 6: one f and one (f, grad f) evaluation, cliff face + LATE;
  vendor     cusolverDnDpotrf / cublasDtrsm / cublasDtrmm / cublasDgemm on the same box beside this library's kernels;
  k1         covariance assembly GB/s against the measured HBM peak.
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
HBM_PEAK_FALLBACK = 6548.8      # GB/s, MEASURED_PEAKS.json of this pool (used if the file is absent on the box)


def flops_per_mll_draw(m):      # n^2 (TRMV) + 2 m_T^2 + 2 m_C^2 + 2 n^2 = 16 m^2
    return 16.0 * m * m


def flops_per_chi_draw(m, nb):  # 8 m^2 + 4 nb m + 2 nb^2
    return 8.0 * m * m + 4.0 * nb * m + 2.0 * nb * nb


def config_dict(world):
    """The SAME dict in both arms (the driver compares them); the reference arm's bounded sample is described in its
    cpu_baseline.sample, not here."""
    return {"workload": "C4: NYC-2015-shaped synthetic district pair, 4000 units/side, Matern-3/2 + Const, "
                        "hypotest_mll null, 20000 draws per GPU per step",
            "m_per_side": M_SIDE, "n": 2 * M_SIDE, "sentinels": NB, "draws_per_gpu_per_step": S_DRAWS,
            "rng": "device Philox-4x32-10 + Box-Muller", "parallelism": f"draw-sharded x{world}",
            "l2": "inputs larger than L2 (factors 768 MB + 6.6 GB of draw buffers per step)"}


def synth(m_side=M_SIDE, nb=NB, seed=1):
    """SURVEY §8d synthetic two-region data (shape-faithful to test/runtests.jl:4-36)."""
    rng = np.random.default_rng(seed)
    n = 2 * m_side
    X = rng.random((2, n))
    rad = np.sqrt(X[0] ** 2 + X[1] ** 2)
    order = np.argsort(rad, kind="stable")
    X, rad = X[:, order], rad[order]
    r = 0.5 * (rad[m_side - 1] + rad[m_side])
    treat = np.arange(n) < m_side
    Y = (-0.8 + np.sin(X[0] * 2.0) + 3.0 * X[1] ** 2 + 0.3 * rng.standard_normal(n) + 0.3 * treat)
    th = np.linspace(0.0, np.pi / 2.0, nb)
    Xb = np.asfortranarray(np.vstack([r * np.cos(th), r * np.sin(th)]))
    return (np.asfortranarray(X[:, :m_side]), np.asfortranarray(X[:, m_side:]), Y[:m_side].copy(), Y[m_side:].copy(), Xb)


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


# ----------------------------------------------------------------------------------------------------------------
# CPU side: the reference's per-draw routine sequence on the oracle (Julia is absent), with the BLAS thread count set
# EXPLICITLY (torchrun exports OMP_NUM_THREADS=1, which silently made the round-1 reference arm single-threaded at N > 1)
# ----------------------------------------------------------------------------------------------------------------
def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class _CpuRef:
    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import geordd_oracle as O
        from threadpoolctl import threadpool_limits
        self.O, self.limits, self.cores = O, threadpool_limits, _host_cores()
        XT, XC, YT, YC, Xb = synth()
        spec = O.KernelSpec(O.MAT32, ELL, SIG2, CONST_S2, 0.0)
        ln = math.log(NOISE_SD)
        t0 = time.time()
        with self.limits(limits=self.cores):
            self.gT, self.gC = O.gp_fit(spec, XT, YT, ln), O.gp_fit(spec, XC, YC, ln)
            self.gN = O.make_null(self.gT, self.gC)
        self.setup_s = time.time() - t0

    def rate(self, draws, threads, rng):
        """mll null draws/s, one draw at a time exactly as src/hypotest_mll.jl:1-30 (dtrmv + 3 dpotrs + dots per draw)."""
        Z = rng.standard_normal((2 * M_SIDE, draws))
        with self.limits(limits=threads):
            t0 = time.time()
            self.O.nsim_logP(self.gT, self.gC, self.gN, Z)
            dt = time.time() - t0
        return draws / dt, dt


def cpu_baseline_block(sample_all=16, sample_one=6):
    ref = _CpuRef()
    rng = np.random.default_rng(2)
    ref.rate(1, ref.cores, rng)                                   # warm-up draw
    r_all, dt_all = ref.rate(sample_all, ref.cores, rng)
    r_one, dt_one = ref.rate(sample_one, 1, rng)
    return {"value": r_all, "unit": UNIT, "cores": ref.cores, "kind": "port",
            "value_1_thread": r_one, "blas_threads": {"all": ref.cores, "one": 1},
            "sample": f"{sample_all} mll null draws at n=8000 on {ref.cores} BLAS threads ({dt_all:.1f} s) and {sample_one} on 1 "
                      f"thread ({dt_one:.1f} s), after {ref.setup_s:.1f} s setup (assemble + 3 dpotrf, excluded); oracle port of "
                      f"src/hypotest_mll.jl:1-19 (dtrmv + 3 dpotrs per draw, SciPy/OpenBLAS; thread count set with threadpoolctl)"}


def run_reference(args, rank, world):
    """--impl reference: the reference CPU path (Julia is absent -> the oracle port of it) on this box's host cores.
    Each step is a bounded sample (24 draws) of the C4 workload; rank 0 alone runs it."""
    if rank != 0:
        return
    ref = _CpuRef()
    sample = 24
    rng = np.random.default_rng(3)
    ref.rate(2, ref.cores, rng)
    vals = []
    for i in range(args.warmup + args.steps):
        r, _ = ref.rate(sample if i >= args.warmup else 2, ref.cores, rng)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean(vals))
    v1, dt1 = ref.rate(6, 1, rng)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref.cores, "kind": "port", "value_1_thread": v1,
                             "blas_threads": {"all": ref.cores, "one": 1},
                             "sample": f"{sample} mll null draws per step at n=8000 (oracle port of src/hypotest_mll.jl:1-19 via "
                                       f"SciPy/OpenBLAS dtrmv + 3 dpotrs per draw) on {ref.cores} BLAS threads set explicitly with "
                                       f"threadpoolctl (OMP_NUM_THREADS of the launcher is ignored); 1-thread rate from 6 draws "
                                       f"({dt1:.1f} s); setup (assemble + 3 dpotrf) {ref.setup_s:.1f} s excluded"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# vendor comparators (with no reference sm_100 kernel, cuSOLVER / cuBLAS on the same box are the bar)
# ----------------------------------------------------------------------------------------------------------------
def vendor_block(torch, dev):
    import ctypes as C
    out = {}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    try:
        torch.backends.cuda.preferred_linalg_library("cusolver")
        for n in (4096, 8192, 16384):
            g = torch.Generator(device=dev).manual_seed(n)
            B = torch.rand(n, 64, device=dev, dtype=torch.float64, generator=g)
            A = B @ B.T
            A.diagonal().add_(float(n))
            ms = timed(lambda: torch.linalg.cholesky(A))
            out[f"cusolver_dpotrf_{n}_ms"] = ms
            out[f"cusolver_dpotrf_{n}_tflops"] = n ** 3 / 3 / (ms * 1e-3) / 1e12
            del A, B
        n, nrhs = 8064, 20032
        g = torch.Generator(device=dev).manual_seed(7)
        L = torch.tril(torch.rand(n, n, device=dev, dtype=torch.float64, generator=g)) / n
        L.diagonal().add_(1.0)
        Bm = torch.rand(n, nrhs, device=dev, dtype=torch.float64, generator=g)
        ms = timed(lambda: torch.linalg.solve_triangular(L, Bm, upper=False))
        out["cublas_dtrsm_8064x20032_ms"] = ms
        out["cublas_dtrsm_8064x20032_tflops"] = n * n * nrhs / (ms * 1e-3) / 1e12
        # cublasDtrmm through ctypes (torch has no trmm): C = L * B, column-major views of the row-major torch tensors
        import glob
        cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cublas", "lib", "libcublas.so.*"))
        cb = C.CDLL(cands[0] if cands else "libcublas.so.12")
        h = C.c_void_p()
        assert cb.cublasCreate_v2(C.byref(h)) == 0
        cb.cublasSetStream_v2(h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        Lc = L.T.contiguous()            # memory = column-major L
        Bc = Bm.T.contiguous()           # memory = column-major B (n x nrhs)
        Cc = torch.empty_like(Bc)
        one = C.c_double(1.0)
        # side left (0), uplo lower (0), op N (0), diag non-unit (0)
        fn = lambda: cb.cublasDtrmm_v2(h, 0, 0, 0, 0, n, nrhs, C.byref(one), C.c_void_p(Lc.data_ptr()), n,
                                       C.c_void_p(Bc.data_ptr()), n, C.c_void_p(Cc.data_ptr()), n)
        ms = timed(fn)
        out["cublas_dtrmm_8064x20032_ms"] = ms
        out["cublas_dtrmm_8064x20032_tflops"] = n * n * nrhs / (ms * 1e-3) / 1e12
        cb.cublasDestroy_v2(h)
        del L, Bm, Lc, Bc, Cc
        n = 8192
        A = torch.rand(n, n, device=dev, dtype=torch.float64)
        B = torch.rand(n, n, device=dev, dtype=torch.float64)
        ms = timed(lambda: A @ B)
        out["cublas_dgemm_8192_ms"] = ms
        out["cublas_dgemm_8192_tflops"] = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        del A, B
        torch.cuda.empty_cache()
    except Exception as err:   # a comparator must never take the bench line down
        out["error"] = repr(err)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="geordd_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--draws", type=int, default=S_DRAWS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    import georddb200
    G = georddb200.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=4))   # a mismatched collective fails fast
    ctx = G.Context(local_rank)
    G.geordd._default_ctx = ctx          # the host drivers below (boot_*, GPRealisations) use this rank's context
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

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    for i in range(max(args.warmup, 3)):
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
    ms = max_over_ranks(ms)
    value = world * S * args.steps / (ms * 1e-3)

    # ---- end to end through the public host API from pinned HOST buffers ----------------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = a
        return v, t

    def pinned_f(a):   # (2 x m) arrays are passed column-major: pinned 1-d storage viewed in Fortran order
        t = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        return v, t

    hXT, _k1 = pinned_f(XT); hXC, _k2 = pinned_f(XC); hYT, _k3 = pinned(YT); hYC, _k4 = pinned(YC)
    out_null = torch.empty(S, dtype=torch.float64, pin_memory=True)
    out_alt = torch.empty(S, dtype=torch.float64, pin_memory=True)
    rdict = {True: G.RegionData(hXT, hYT, np.empty((0, M_SIDE))), False: G.RegionData(hXC, hYC, np.empty((0, M_SIDE)))}

    def e2e_step(i):
        gpr = G.GPRealisations(rdict, k, ln, groupKeys=[True, False], ctx=ctx)   # H2D coords + y; both fits in one call
        a, b = gpr[True], gpr[False]
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
    e2e_step(1)
    n_e2e = max(10, args.steps)
    e2e_ms = []
    sampler2 = ClockSampler(local_rank)
    sampler2.start()
    for i in range(n_e2e):
        barrier()
        t_wall0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        e2e_step(2 + i)
        f1.record(stream)
        torch.cuda.synchronize()
        e2e_ms.append(max_over_ranks(max(f0.elapsed_time(f1), (time.perf_counter() - t_wall0) * 1e3)))   # host work counts too
    clocks_e2e = sampler2.stop()
    ms_e2e = float(np.median(e2e_ms))
    e2e_val = world * S / (ms_e2e * 1e-3)
    h2d = 8 * (XT.size + XC.size + YT.size + YC.size) * 2            # T, C and the pooled null each upload X and y
    d2h = 8 * (2 * S + 2 * M_SIDE * 2) + 8 * 8                       # statistics + alphas + tallies

    # ---- the other named shapes (all ranks take part where the draws are sharded) ---------------------------
    extra = {}
    if not args.no_extra:
        dmma_peak = DMMA_PEAK_UNTHROTTLED

        def wall(fn, reps=3):
            fn()
            best = 1e30
            for _ in range(reps):
                barrier(); t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                best = min(best, max_over_ranks((time.perf_counter() - t0) * 1e3))
            return best

        # C4 strong scaling: 20 000 draws IN TOTAL over the ranks, per-call wall time of boot_mlltest (incl. the pooled fit)
        ms_c4s = wall(lambda: G.boot_mlltest(gT, gC, S_DRAWS, seed=5))
        extra["c4_strong"] = {"what": "boot_mlltest(gpT, gpC, 20000): make_null (pooled fit, n=8000) + 20000 draws in total "
                                      "sharded over the ranks + tally all-reduce, wall time per call",
                              "draws_total": S_DRAWS, "ms_per_call": ms_c4s, "sims_per_sec": S_DRAWS / (ms_c4s * 1e-3),
                              "flops": flops_per_mll_draw(M_SIDE) * S_DRAWS + (2 * M_SIDE) ** 3 / 3.0 * world,
                              "tflops_per_gpu": (flops_per_mll_draw(M_SIDE) * S_DRAWS / world + (2 * M_SIDE) ** 3 / 3.0) / (ms_c4s * 1e-3) / 1e12}
        extra["c4_strong"]["frac_of_dmma_peak"] = extra["c4_strong"]["tflops_per_gpu"] / dmma_peak
        # C3: Mississippi-shaped sweep
        m3, nb3, S3 = 1000, 200, 10000
        X3T, X3C, Y3T, Y3C, X3b = synth(m3, nb3, seed=3)
        g3 = G.GPRealisations({True: G.RegionData(X3T, Y3T, np.empty((0, m3))), False: G.RegionData(X3C, Y3C, np.empty((0, m3)))},
                              k, ln, groupKeys=[True, False], ctx=ctx)
        ms_chi = wall(lambda: G.boot_chi2test(g3[True], g3[False], X3b, S3, seed=6), reps=9)
        ms_inv = wall(lambda: G.boot_invvar(g3[True], g3[False], X3b, S3, seed=6), reps=9)
        f3 = flops_per_chi_draw(m3, nb3) * S3
        extra["c3"] = {"what": "1000 units/side, 200 sentinels, 10000 draws in total sharded over the ranks; wall time per call "
                               "incl. make_null, cliff face and the observed statistic",
                       "boot_chi2test_ms": ms_chi, "boot_invvar_ms": ms_inv, "chi2_sims_per_sec": S3 / (ms_chi * 1e-3),
                       "invvar_sims_per_sec": S3 / (ms_inv * 1e-3), "flops_draws": f3,
                       "tflops_per_gpu_chi2": f3 / world / (ms_chi * 1e-3) / 1e12,
                       "frac_of_dmma_peak_chi2": f3 / world / (ms_chi * 1e-3) / 1e12 / dmma_peak}
        g3[True].close(); g3[False].close()
        # C5: power study shape (16 000 units): null reference vectors (1000 draws x 3 statistics), then 500 power draws
        m5, S5n, S5p = 8000, 1000, 500
        X5T, X5C, Y5T, Y5C, X5b = synth(m5, 100, seed=1)
        k5 = G.SEIso(math.log(ELL), 0.0) + G.Const(0.5 * math.log(CONST_S2))
        t0 = time.perf_counter()
        g5 = G.GPRealisations({True: G.RegionData(X5T, Y5T, np.empty((0, m5))), False: G.RegionData(X5C, Y5C, np.empty((0, m5)))},
                              k5, ln, groupKeys=[True, False], ctx=ctx)
        n5 = G.NullGPE(g5[True], g5[False], X5b, 0.0)
        torch.cuda.synchronize()
        fit5_ms = (time.perf_counter() - t0) * 1e3

        def c5_null():
            chi = G.nsim_chi(g5[True], g5[False], n5, X5b, S5n, seed=7, sharded=True)
            _, _, mn, ma = G.nsim_logP(g5[True], g5[False], n5, S5n, seed=7, return_count=True, sharded=True)
            pinv = G.nsim_invvar_pval_uncalib(g5[True], g5[False], X5b, S5n, seed=7, gpNull=n5, sharded=True)
            return chi, ma - mn, pinv
        chi5, dm5, pinv5 = c5_null()
        ms_c5n = wall(c5_null, reps=2)
        ms_c5p = wall(lambda: G.nsim_power(g5[True], g5[False], 0.3, X5b, chi5, dm5, pinv5, S5p, seed=8, sharded=True, gpNull=n5), reps=2)
        f5p = (flops_per_mll_draw(m5) + 4.0 * 100 * m5 + 4.0 * 100 * 100) * S5p
        extra["c5"] = {"what": "16000 units (8000/side), SEIso + Const, 100 sentinels: fits (T, C batched; pooled null n=16000), "
                               "1000 null draws x (chi2, mll, invvar) and nsim_power with 500 alternative draws (5 p-values each), "
                               "draws sharded over the ranks, vectors all-gathered",
                       "fits_ms": fit5_ms, "null_vectors_ms": ms_c5n, "nsim_power_ms": ms_c5p,
                       "power_draws_per_sec": S5p / (ms_c5p * 1e-3), "flops_power": f5p,
                       "tflops_per_gpu_power": f5p / world / (ms_c5p * 1e-3) / 1e12,
                       "frac_of_dmma_peak_power": f5p / world / (ms_c5p * 1e-3) / 1e12 / dmma_peak}
        n5.close(); g5[True].close(); g5[False].close()

    if world > 1:
        # every collective of the run is behind us: leave the process group BEFORE rank 0's private measurements, so that no
        # rank sits in an NCCL barrier (with its watchdog) while rank 0 times single-GPU shapes and the CPU baseline
        dist.barrier()
        dist.destroy_process_group()

    if rank == 0:
        # ---- rank 0 only: FP64 tensor-pipe ceiling, Cholesky, cliff-face latency, C2, K1, vendor ----------------
        # (the host drivers shard and all-reduce automatically under torch.distributed; nothing in this block may enter a
        #  collective, the other ranks are not here)
        _local = G.parallel.local_only()
        _local.__enter__()
        dmma_live = ctx.dmma_peak_tflops()
        # The register-resident DMMA loop draws more power than any real kernel; on a power-capped box it measures
        # below what the solve kernels themselves sustain.  The roofline denominator is therefore the larger of the
        # live probe and the unthrottled probe value measured on this pool in round 1 (37.07 TFLOP/s =
        # 148 SMs x 128 FP64 tensor flop/clk x 1.965 GHz to 0.4 %; profiles/r01_bench_first.json).
        dmma_peak = max(dmma_live, DMMA_PEAK_UNTHROTTLED)
        if not args.no_extra:
            potrf = {}
            for n in (4096, 8192, 16384):
                pms, _ = ctx.bench_potrf(n, 3)
                potrf[str(n)] = {"ms": pms, "tflops": n ** 3 / 3 / (pms * 1e-3) / 1e12, "frac_of_dmma_peak": n ** 3 / 3 / (pms * 1e-3) / 1e12 / dmma_peak}
            extra["potrf"] = potrf

            def host_ms(fn, reps=5):
                fn(); ctx.sync()
                ts = []
                for _ in range(reps):
                    t0 = time.perf_counter(); fn(); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
                return float(np.median(ts))

            def fit_and_cliff():
                gpr = G.GPRealisations(rdict, k, ln, groupKeys=[True, False], ctx=ctx)
                G.cliff_face(gpr[True], gpr[False], Xb)
                gpr[True].close(); gpr[False].close()
            cliff_fit_ms = host_ms(fit_and_cliff)
            cliff_only_ms = host_ms(lambda: G.cliff_face(gT, gC, Xb))
            f_fit = 2 * (M_SIDE ** 3 / 3.0 + 2.0 * M_SIDE ** 2) + 2 * (M_SIDE ** 2 * NB + M_SIDE * NB ** 2 + 2.0 * M_SIDE * NB)
            extra["cliff_face"] = {"what": "GPRealisations(T, C) from host buffers (assemble, Cholesky, alpha, mll; both regions in one "
                                           "call) + cliff_face at 100 sentinels, host wall time per call",
                                   "fit_plus_cliff_ms_4k_per_side": cliff_fit_ms, "cliff_only_ms_4k_per_side": cliff_only_ms,
                                   "flops": f_fit, "tflops": f_fit / (cliff_fit_ms * 1e-3) / 1e12,
                                   "frac_of_dmma_peak": f_fit / (cliff_fit_ms * 1e-3) / 1e12 / dmma_peak}
            cnt = C.c_longlong()
            for rep in range(2):   # first call allocates the sentinel-system buffers of the handle
                ctx.sync(); t0 = time.perf_counter()
                ctx.check(lib.geordd_null_chi2(gN._h, S, 1, 10 ** 9, None, 1e300, None, C.byref(cnt)))
                ctx.sync(); chi_rate = S / (time.perf_counter() - t0)
            extra["chi2_sims_per_sec_1gpu"] = chi_rate
            extra["chi2_frac_of_dmma_peak"] = chi_rate * flops_per_chi_draw(M_SIDE, NB) / 1e12 / dmma_peak
            # opt-in forward-only evaluation of the same mll statistic (|L^-1 (y-m)|^2 and z'z; not the headline path)
            ctx.set_mll_forward_only(True)
            try:
                for rep in range(2):
                    ctx.sync(); t0 = time.perf_counter()
                    ctx.check(lib.geordd_null_mll(gN._h, S, 1, 10 ** 9, None, d_obs, None, None, C.byref(cnt), None, None))
                    ctx.sync(); extra["mll_forward_only_sims_per_sec_1gpu"] = S / (time.perf_counter() - t0)
            finally:
                ctx.set_mll_forward_only(False)
            # C2: joint fit with covariates, 2 000 units/side, p = 6
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            m2 = 2000
            rng2 = np.random.default_rng(2)
            X2T, X2C, Y2T, Y2C, X2b = synth(m2, 100, seed=2)
            D2 = np.asfortranarray(np.vstack([np.ones(2 * m2), rng2.standard_normal(2 * m2), rng2.standard_normal(2 * m2),
                                              (rng2.integers(0, 3, 2 * m2)[None, :] == np.arange(3)[:, None]).astype(float)]))
            rd2 = {True: G.RegionData(X2T, Y2T, np.asfortranarray(D2[:, :m2])), False: G.RegionData(X2C, Y2C, np.asfortranarray(D2[:, m2:]))}
            mg = G.MultiGPCovars(rd2, G.SEIso(math.log(ELL), 0.0) + G.Const(0.5 * math.log(CONST_S2)), G.LinIso(math.log(10.0)), ln,
                                 groupKeys=[True, False], ctx=ctx)
            f_ms = host_ms(lambda: G.update_mll(mg))
            g_ms = host_ms(lambda: G.update_mll_and_dmll(mg, noise=True, domean=False, kern=True, beta=True))
            n2 = 2 * m2

            def c2_late():
                beta = G.postmean_beta(mg)
                res = {kk: G.residuals_data(rd, beta) for kk, rd in rd2.items()}
                gpr = G.GPRealisations(res, mg.kernel, mg.logNoise, groupKeys=[True, False], ctx=ctx)
                mu, Sg = G.cliff_face(gpr[True], gpr[False], X2b)
                t = G.inverse_variance(mu, Sg)
                gpr[True].close(); gpr[False].close()
                return t
            late_ms = host_ms(c2_late, reps=3)
            extra["c2"] = {"what": "MultiGPCovars at n = 4000, p = 6: update_mll! (f), update_mll_and_dmll! (f and all 5 gradient "
                                   "entries), then postmean_beta + residual GPRealisations + cliff face + inverse-variance LATE; host "
                                   "wall time per call, one ABI call per evaluation",
                           "f_ms": f_ms, "f_and_grad_ms": g_ms, "postmean_to_late_ms": late_ms,
                           "f_flops": n2 ** 3 / 3.0, "f_tflops": n2 ** 3 / 3.0 / (f_ms * 1e-3) / 1e12,
                           "grad_flops": n2 ** 3 / 3.0 + 2.0 * n2 ** 3, "grad_tflops": (n2 ** 3 / 3.0 + 2.0 * n2 ** 3) / (g_ms * 1e-3) / 1e12,
                           "grad_frac_of_dmma_peak": (n2 ** 3 / 3.0 + 2.0 * n2 ** 3) / (g_ms * 1e-3) / 1e12 / dmma_peak}
            mg.close()
            # f2: optimize!(mgpcv) on the reference's own example (notebooks/SimulatedExample.ipynb cell 10: n = 300, the 604-column
            # full-dummy design, 167 f + 99 grad evaluations in 7.95 s on the author's CPU, optimum 123.1640)
            try:
                fx = np.load(os.path.join(ROOT, "tests", "golden", "simulated_example.npz"))
                dummy = lambda v: (np.asarray(v)[None, :] == np.sort(np.unique(v))[:, None]).astype(np.float64)
                Dn = np.asfortranarray(np.vstack([np.ones((1, 300)), dummy(fx["covarA"]), dummy(fx["covarB"]), dummy(fx["categ"])]))
                tr = fx["treat"]
                rdn = {False: G.RegionData(np.asfortranarray(fx["X"][:, ~tr]), fx["Y"][~tr].copy(), np.asfortranarray(Dn[:, ~tr])),
                       True: G.RegionData(np.asfortranarray(fx["X"][:, tr]), fx["Y"][tr].copy(), np.asfortranarray(Dn[:, tr]))}
                best = None
                for _ in range(2):
                    mgn = G.MultiGPCovars(rdn, G.SEIso(math.log(0.5), 0.0) + G.fix(G.Const(-12.0)), G.LinIso(0.0), 1.0, groupKeys=[False, True], ctx=ctx)
                    t0 = time.perf_counter()
                    res = G.optimize(mgn, noise=True, kern=True, domean=False, beta=True, method="BFGS", options=dict(gtol=1e-7, maxiter=500))
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                    fmin = -mgn.mll
                    mgn.close()
                extra["f2_optimize"] = {"what": "optimize!(MultiGPCovars) on the data of notebooks/SimulatedExample.ipynb (n = 300, p = 604; replayed "
                                                "bit for bit), from the notebook's initial point; every evaluation is one ABI call; host optimiser = SciPy BFGS",
                                        "seconds": best, "evaluations": int(res.nfev), "minimum": fmin, "notebook_minimum": 123.1640,
                                        "notebook_seconds": 7.95, "notebook_evaluations": "167 f + 99 grad"}
            except Exception as err:
                extra["f2_optimize"] = {"error": repr(err)}
            # f3: placebo sweep of the NYC notebook shape (notebooks/NYC_analysis-2015.ipynb:1452-1512: 90 angles x 1000 draws; the
            # notebook's @time outputs: mll 73.5 / 917.1 s, chi2 36.6 / 430.0 s for its two districts)
            try:
                rngp = np.random.default_rng(7)
                npl = 1000
                Xp = np.asfortranarray(rngp.random((2, npl)))
                Yp = np.sin(2.0 * Xp[0]) + 3.0 * Xp[1] ** 2 + 0.3 * rngp.standard_normal(npl)
                kp = G.Mat12Iso(math.log(ELL), 0.0) + G.Const(math.log(5.0))
                angles = [float(a) for a in range(1, 180, 2)]
                f3 = {"what": "placebo_mll / placebo_chi for angle in 1:2:180 (90 angles) x 1000 draws on a 1000-unit district, Matern-1/2 + Const; "
                              "4 host threads, each with its own context, run independent (split, 2 fits, pooled fit, test) problems side by side",
                      "notebook_seconds": {"mll": [73.5, 917.1], "chi2": [36.6, 430.0]}}
                for which in ("mll", "chi"):
                    G.placebo_sweep(which, angles[:4], Xp, Yp, kp, G.MeanZero(), ln, 1000, seed=1, workers=4)
                    t0 = time.perf_counter()
                    G.placebo_sweep(which, angles, Xp, Yp, kp, G.MeanZero(), ln, 1000, seed=1, workers=4)
                    f3[f"placebo_{which}_seconds"] = time.perf_counter() - t0
                extra["f3_placebo_sweep"] = f3
            except Exception as err:
                extra["f3_placebo_sweep"] = {"error": repr(err)}
            # K1: covariance assembly against the HBM peak (8 n^2 bytes written + 16 n read)
            try:
                hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                hbm_src = "MEASURED_PEAKS.json"
            except Exception:
                hbm, hbm_src = HBM_PEAK_FALLBACK, "fallback 6548.8 GB/s (MEASURED_PEAKS.json absent on the box)"
            ctx.profile(True)
            for _ in range(5):
                tmp = G.NullGPE(gT, gC)       # one pooled fit = one symmetric assembly at n = 8000 (padded 8064)
                tmp.close()
            cov_ms, cov_n = ctx.profile_get("cov")
            ctx.profile(False)
            npad = 8064
            cov_bytes = 8.0 * npad * npad / 2 + 8.0 * npad * 128 / 2      # lower tiles only (the factorisation reads nothing else)
            extra["k1"] = {"kernel": "cov_kernel (symmetric assembly, lower tiles, n = 8000 padded to 8064, Matern-3/2 + Const)",
                           "avg_launch_ms": cov_ms / max(cov_n, 1), "bytes_per_launch": cov_bytes,
                           "gbs": cov_bytes / (cov_ms / max(cov_n, 1) * 1e-3) / 1e9, "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src,
                           "frac_of_hbm_peak": cov_bytes / (cov_ms / max(cov_n, 1) * 1e-3) / 1e9 / hbm}
            extra["vendor"] = vendor_block(torch, dev)
            extra["dmma_peak_tflops"] = dmma_peak
        extra["profile_ms"] = {k_: v_[0] for k_, v_ in prof.items()}
        extra["e2e_ms_all_steps"] = e2e_ms

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
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
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
                    "also": {"kernel": "solve_dataflow_kernel<TRMM> (prior_rand with the Philox draw fused into its producer, 4 m^2 flop/draw)",
                             "achieved": trmm_flops * trmm_n / (trmm_ms * 1e-3) / 1e12, "avg_launch_ms": trmm_ms / max(trmm_n, 1),
                             "share_of_step": trmm_ms / ms},
                    "all_dmma_launches": {"achieved": flops_per_mll_draw(M_SIDE) * S * args.steps / (dmma_ms * 1e-3) / 1e12,
                                          "share_of_step": dmma_ms / ms}}
        cpu = None if args.no_cpu_baseline else cpu_baseline_block()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(world),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "what": "per step: GPRealisations(T, C) + make_null from pinned host X/y + 20000 mll draws + D2H of both "
                                "statistic vectors (boot_mlltest through the host API); median over the steps", "steps": n_e2e,
                        "ms_per_step_median": ms_e2e, "ms_per_step_min": float(np.min(e2e_ms)), "ms_per_step_max": float(np.max(e2e_ms)),
                        "clocks": clocks_e2e},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "cpu_baseline": cpu,
                "extra": extra}
        print(json.dumps(line))
        _local.__exit__(None, None, None)


if __name__ == "__main__":
    main()
