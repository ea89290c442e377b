"""Multi-GPU check of the boot_* drivers (run under torchrun on N GPUs): the p-values computed with the draws sharded
over the ranks equal, bit for bit, the single-process values (draws are keyed by (seed, index)).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/gpu_check_multigpu.py"""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist
import georddb200, geordd_oracle as O
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
G = georddb200.load(); ctx = G.Context(lr)
d = O.synth_two_region(400, 50, seed=1, tau=0.1)
k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
gT, gC = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, ln, ctx=ctx), G.GPE(d["XC"], d["YC"], G.MeanZero(), k, ln, ctx=ctx)
S = 10007
single = (G.boot_chi2test(gT, gC, d["Xb"], S, seed=7), G.boot_mlltest(gT, gC, S, seed=7), G.boot_invvar(gT, gC, d["Xb"], S, seed=7))
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sharded = (G.boot_chi2test(gT, gC, d["Xb"], S, seed=7), G.boot_mlltest(gT, gC, S, seed=7), G.boot_invvar(gT, gC, d["Xb"], S, seed=7))
ang = G.placebo_sweep("invvar", [10.0, 40.0, 70.0, 100.0, 130.0], d["XT"], d["YT"], k, G.MeanZero(), ln)
dist.barrier()
assert single == sharded, (single, sharded)
assert ang.shape == (5,) and np.all((ang >= 0) & (ang <= 1))
if rank == 0:
    print(f"world {world}: p-values (chi2, mll, invvar) single == sharded == {sharded}; placebo sweep {np.round(ang, 4)}")
    print("MULTIGPU OK")
dist.destroy_process_group()
