"""CPU tests: the C-ABI library loads and exports every symbol include/geordd_b200.h declares (no
compute without a GPU), the host-side mirror's pure-Python logic, and the multi-GPU tally reduction
over gloo with world_size 2."""
import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(G):
    hdr = open(os.path.join(ROOT, "include", "geordd_b200.h")).read()
    declared = set(re.findall(r"GEORDD_API\s+(?:const\s+char\s*\*|int|long long|geordd_\w+\s*\*)\s*(geordd_\w+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = G.capi.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(G.capi.EXPORTED_SYMBOLS), declared ^ set(G.capi.EXPORTED_SYMBOLS)
    assert lib.geordd_version() == 100


def test_no_cpu_fallback_without_gpu(G):
    """Product code must fail loudly when no device is usable (and never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(G.capi.GeorddError):
        G.Context(0)
    src = ""
    for f in ("geordd.py", "capi.py", "parallel.py", "__init__.py"):
        src += open(os.path.join(ROOT, "geordd.jl_b200", f)).read()
    assert "geordd_oracle" not in src and "import oracle" not in src


def test_kernel_pattern_matching(G):
    k = G.SEIso(math.log(0.3), 0.0) + G.Const(math.log(20.0))
    s = G.kernel_spec(k, G.MeanZero())
    assert (s.kind, s.const_sigma2, s.mean_const) == (0, pytest.approx(400.0), 0.0)
    assert s.ell == pytest.approx(0.3) and s.sigma2 == pytest.approx(1.0)
    s = G.kernel_spec(G.fix(G.Const(0.0)) + G.Mat32Iso(math.log(2.0), math.log(3.0)), G.MeanConst(1.5))
    assert (s.kind, s.ell, s.sigma2, s.const_sigma2, s.mean_const) == (2, pytest.approx(2.0), pytest.approx(9.0), 1.0, 1.5)
    assert G.kernel_spec(G.Mat12Iso(0.0, 0.0)).kind == 1 and G.kernel_spec(G.Mat52Iso(0.0, 0.0)).kind == 3
    with pytest.raises(ValueError):
        G.kernel_spec(G.LinIso(0.0))                       # ArgumentError, no fallback
    with pytest.raises(ValueError):
        G.kernel_spec(G.SEIso(0, 0) + G.SEIso(0, 0))
    assert G.get_params(k) == [math.log(0.3), 0.0, math.log(20.0)]
    G.set_params(k, [0.1, 0.2, 0.3])
    assert (k.kleft.ll, k.kleft.lsigma, k.kright.lsigma) == (0.1, 0.2, 0.3)
    kf = G.SEIso(0.0, 0.0) + G.fix(G.Const(1.0))
    assert G.get_params(kf) == [0.0, 0.0]


def test_placebo_geometry(G):
    rng = np.random.default_rng(0)
    X = rng.random((2, 401))
    for angle in (10.0, 30.0, 80.0, 135.0, 271.0):
        shift = G.shift_for_even_split(angle, X)
        left = G.left_points(angle, shift, X)
        assert abs(int(left.sum()) - 200.5) <= 1.0
        Xb = G.placebo_sentinels(angle, shift, X, 100)
        assert Xb.shape == (2, 100)
        assert Xb[0].min() >= X[0].min() - 1e-9 and Xb[0].max() <= X[0].max() + 1e-9
        assert Xb[1].min() >= X[1].min() - 1e-9 and Xb[1].max() <= X[1].max() + 1e-9
        # sentinels are collinear with slope tan(angle) and separate the two halves
        dx, dy = Xb[0, -1] - Xb[0, 0], Xb[1, -1] - Xb[1, 0]
        assert abs(math.atan2(dy, dx) % math.pi - math.radians(angle) % math.pi) < 1e-9
        cross = dx * (X[1] - Xb[1, 0]) - dy * (X[0] - Xb[0, 0])
        side = cross > 0
        assert side.sum() in (int(left.sum()), 401 - int(left.sum()))


def test_placebo_geometry_matches_oracle_and_exact_degree_trig(G, O):
    """src/placebo_geometry.jl uses tand / cosd / sind: exact at the special angles the notebook sweep `1:2:180` hits
    (tand(45) == 1 puts 45 deg in the `else` branch of left_points; cosd(90) == 0 makes `direction` fall back to 1).
    The host mirror (libm on the degree-reduced argument) against the oracle (mpmath on the exact degree value)."""
    assert G.geordd._tand(45.0) == 1.0 == O.tand(45.0) and G.geordd._tand(135.0) == -1.0 == O.tand(135.0)
    assert G.geordd._cosd(90.0) == 0.0 == O.cosd(90.0) and G.geordd._cosd(270.0) == 0.0 and G.geordd._sind(180.0) == 0.0
    assert G.geordd._tand(90.0) == math.inf == O.tand(90.0) and G.geordd._sind(30.0) == pytest.approx(0.5, abs=1e-16)
    for a in np.arange(-30.0, 400.0, 7.5):
        assert G.geordd._sind(a) == pytest.approx(O.sind(a), abs=2e-16) and G.geordd._cosd(a) == pytest.approx(O.cosd(a), abs=2e-16)
    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.random((2, 260)) * np.array([[3.0], [1.0]]) + np.array([[10.0], [-2.0]]))
    for angle in [float(a) for a in range(1, 180, 2)] + [0.0, 90.0, 180.0, 44.99]:
        sg, so = G.shift_for_even_split(angle, X), O.shift_for_even_split(angle, X)
        assert sg == so, (angle, sg, so)
        assert np.array_equal(G.left_points(angle, sg, X), O.left_points(angle, so, X))
        bg, bo = G.placebo_sentinels(angle, sg, X, 100), O.placebo_sentinels(angle, so, X, 100)
        assert np.max(np.abs(bg - bo)) < 1e-14


def test_residuals_and_normal(G):
    rd = G.RegionData(np.zeros((2, 3)), np.array([1.0, 2.0, 3.0]), np.array([[1.0, 1.0, 1.0], [0.0, 1.0, 2.0]]))
    r = G.residuals_data(rd, [0.5, 1.0])
    assert np.allclose(r.y, [0.5, 0.5, 0.5]) and r.D.shape == (0, 3)
    nrm = G.Normal(0.3, 2.0)
    assert nrm.cdf(0.3) == pytest.approx(0.5) and nrm.cdf(-1) + nrm.ccdf(-1) == pytest.approx(1.0)
    assert nrm.quantile(0.5) == pytest.approx(0.3) and nrm.quantile(0.975) == pytest.approx(0.3 + 2 * 1.959964, abs=1e-5)


def test_shard_draws_partition(G):
    P = G.parallel
    for nsim in (0, 1, 7, 64, 20000, 20001):
        for world in (1, 2, 3, 8):
            spans = [P.shard_draws(nsim, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == nsim
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        P.shard_draws(10, 2, 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import math
import numpy as np, torch.distributed as dist
import georddb200
G = georddb200.load()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
nsim = 1001
stats = np.sin(np.arange(nsim) * 0.37) * 3.0           # stand-in for per-draw statistics keyed by draw index
lo, cnt = G.parallel.shard_draws(nsim, rank, world)
local = stats[lo:lo + cnt]
tally = G.parallel.allreduce_tallies([int((local > 1.0).sum()), int((local < -2.0).sum())])
full = G.parallel.gather_stats(local, nsim)
assert tally == [int((stats > 1.0).sum()), int((stats < -2.0).sum())], tally
assert np.array_equal(full, stats)
angles = list(range(1, 180, 7))                          # the coarse axis: independent problems dealt to the ranks
calls = []
res = G.parallel.map_sharded(lambda a: (calls.append(a), math.cos(a * 0.1))[1], angles)
assert np.array_equal(res, np.cos(np.array(angles) * 0.1)) and len(calls) in (len(angles) // 2, len(angles) - len(angles) // 2)
# vector-returning nsim_* forms (sharded=True): shard bookkeeping + all-gather; host Z columns follow the shard
Zfull = np.arange(3 * 7, dtype=float).reshape(3, 7)
local_n, kw, gather = G.geordd._vector_shard(7, dict(seed=9, draw0=100, Z=Zfull))
lo7, cnt7 = G.parallel.shard_draws(7, rank, world)
assert local_n == cnt7 and kw["draw0"] == 100 + lo7 and np.array_equal(kw["Z"], Zfull[:, lo7:lo7 + cnt7]) and kw["seed"] == 9
assert np.array_equal(gather(kw["Z"][0] * 2.0), Zfull[0] * 2.0)
# a rank whose shard is EMPTY (world > nsim) still joins the tally all-reduce (boot_* drivers; round-1 deadlock)
local_n, kw, reduce_ = G.geordd._draw_shard(1, dict(seed=1))
assert local_n == (0 if rank == 0 else 1)
assert reduce_(5 * local_n) == 5
# rank-0-only work must not enter collectives: local_only() switches the automatic sharding off
if rank == 0:
    with G.parallel.local_only():
        assert not G.parallel.active()
        n_loc, kw2, red2 = G.geordd._draw_shard(10, dict(seed=1))
        assert n_loc == 10 and red2(7) == 7
        assert np.array_equal(G.parallel.map_sharded(lambda a: a * 2.0, [1.0, 2.0, 3.0]), [2.0, 4.0, 6.0])
assert G.parallel.active()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_tally_allreduce_gloo(tmp_path):
    """world_size-2 gloo run of the draw sharding + the single int64 all-reduce."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=29613))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def _build_c_driver(tmp_path):
    exe = str(tmp_path / "c_driver")
    libdir = os.path.join(ROOT, "geordd.jl_b200", "lib")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_driver.c"), "-L" + libdir, "-lgeordd_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_links(G, tmp_path):
    """include/geordd_b200.h compiles as strict C99 and a plain-C consumer links against the library (the boundary
    carries no C++ / torch / Python types).  Without a GPU the driver exits with its "no context" code."""
    G.capi.load_library()
    exe = _build_c_driver(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no usable GPU" in r.stderr
    # the multi-GPU driver (group layer, NCCL bound at run time) is plain C99 as well
    libdir = os.path.join(ROOT, "geordd.jl_b200", "lib")
    exe2 = str(tmp_path / "c_driver_multi")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_driver_multi.c"), "-L" + libdir, "-lgeordd_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe2]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([exe2, "1"], capture_output=True, text=True)
        assert r.returncode == 2
