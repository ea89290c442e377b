"""Per-phase wall times of the end-to-end step of bench.py (fit T,C / pooled null fit / draws / tally all-reduce) on every rank."""
import ctypes as C, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import georddb200, bench
G = georddb200.load()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = G.Context(lr); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
XT, XC, YT, YC, Xb = bench.synth()
k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
rd = {True: G.RegionData(XT, YT, np.empty((0, 4000))), False: G.RegionData(XC, YC, np.empty((0, 4000)))}
S = 20000
tally = torch.zeros(1, dtype=torch.int64, device=dev)
rows = []
for i in range(12):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gpr = G.GPRealisations(rd, k, ln, groupKeys=[True, False], ctx=ctx); a, b = gpr[True], gpr[False]
    t1 = time.perf_counter()
    nn = G.NullGPE(a, b)
    t2 = time.perf_counter()
    cnt = C.c_longlong()
    ctx.check(ctx.lib.geordd_null_mll(nn._h, S, 1, (i * world + rank) * S, None, a.mll + b.mll - nn.mll, None, None, C.byref(cnt), None, None))
    t3 = time.perf_counter()
    if world > 1:
        tally[0] = cnt.value; dist.all_reduce(tally); tally.item()
    t4 = time.perf_counter()
    nn.close(); a.close(); b.close()
    t5 = time.perf_counter()
    rows.append([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4)])
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        print(f"rank {rank}: fitTC  null  draws  allreduce  close  (ms)")
        for row in rows:
            print("   " + "  ".join(f"{v:7.2f}" for v in row), f"  total {sum(row):7.2f}")
        sys.stdout.flush()
if world > 1:
    dist.destroy_process_group()
