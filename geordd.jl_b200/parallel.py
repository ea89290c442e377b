"""Multi-GPU host logic: the ONLY parallel structure of the hot path is the embarrassingly parallel
Monte-Carlo draw loop (`[sim...() for _ in 1:nsim]`, src/hypotest_chi2.jl:49-50, src/hypotest_mll.jl:27-28,
src/hypotest_invvar.jl:60-61, src/power.jl:67-71).  One process per GPU; draw indices are sharded
contiguously; each draw depends only on (seed, draw index), so results are invariant to the GPU count;
the exceedance tallies (a few int64) are summed with ONE all-reduce (NCCL over NVLink on GPUs, gloo in the
CPU tests).  Single fits and cliff faces stay on one GPU (BASELINE north_star)."""
from __future__ import annotations

import contextlib
from typing import Sequence, Tuple

_LOCAL_ONLY = 0


@contextlib.contextmanager
def local_only():
    """Inside this block the host drivers behave as in a single process even if torch.distributed is initialised: no draw
    sharding, no collectives.  For work that only SOME ranks execute (rank-0-only sections of a benchmark): a collective
    entered by a subset of the ranks never returns."""
    global _LOCAL_ONLY
    _LOCAL_ONLY += 1
    try:
        yield
    finally:
        _LOCAL_ONLY -= 1


def active():
    """True if this process is one rank of an initialised torch.distributed job of more than one rank (and not inside
    local_only())."""
    import sys
    if _LOCAL_ONLY:
        return False
    dist = sys.modules.get("torch.distributed")
    return dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_draws(nsim: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [draw0, draw0 + count) of draw indices 0..nsim-1 for `rank`."""
    if world <= 0 or not (0 <= rank < world) or nsim < 0:
        raise ValueError("bad shard request")
    lo = (nsim * rank) // world
    hi = (nsim * (rank + 1)) // world
    return lo, hi - lo


def _auto_device(device):
    """NCCL moves device tensors only: default to this rank's current CUDA device under the nccl backend."""
    if device is not None:
        return device
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return None


def allreduce_tallies(counts: Sequence[int], device=None):
    """Sum integer tallies over all ranks (torch.distributed; backend chosen by the caller's init)."""
    import torch
    import torch.distributed as dist
    if not active():
        return [int(v) for v in counts]
    t = torch.tensor(list(counts), dtype=torch.int64, device=_auto_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.cpu().tolist()]


def gather_stats(local, nsim: int, device=None):
    """All-gather the per-draw statistics (8*S bytes) when the caller wants the full vector (`nsim_*`)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not active():
        return np.asarray(local)
    world = dist.get_world_size()
    device = _auto_device(device)
    sizes = [shard_draws(nsim, r, world)[1] for r in range(world)]
    mx = max(sizes)
    buf = torch.zeros(mx, dtype=torch.float64, device=device)
    buf[: len(local)] = torch.as_tensor(np.asarray(local), dtype=torch.float64).to(buf.device)
    out = [torch.zeros(mx, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)])


def map_sharded(fn, items, device=None, mapper=None):
    """The second, coarser parallel axis of SURVEY §8e: independent problems (placebo angles, adjacent region pairs --
    notebooks/NYC_analysis-2015.ipynb:1452-1555, src/regiondata.jl:64-68) are dealt contiguously to the ranks, each
    rank evaluates `fn(item)` (a float) for its share on its own GPU with no data-path collective, and the results
    are assembled with one all-gather.  Single process: plain map.  `mapper(fn, items)` replaces the in-process loop
    (e.g. a thread pool whose workers own separate GPU contexts)."""
    import numpy as np
    import torch.distributed as dist
    items = list(items)
    run = mapper if mapper is not None else (lambda f, its: [f(it) for it in its])
    if not active():
        return np.array([float(v) for v in run(fn, items)])
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, cnt = shard_draws(len(items), rank, world)
    local = np.array([float(v) for v in run(fn, items[lo:lo + cnt])])
    return gather_stats(local, len(items), device=device)
