"""GPU tests of the multi-device group layer of the C ABI (geordd_group_*): fits replicated, null draws
sharded by index, tallies all-reduced with NCCL inside the library.  Every result must equal the single-GPU path bit for
bit.  The 1-device group runs on any box; the 2-device cases (and the plain-C driver on 2 GPUs) need `gpurun --gpus 2`."""
import math
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [1, 2])
def test_group_matches_single_gpu_bit_for_bit(G, O, ctx, ndev):
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    d = O.synth_two_region(300, 40, seed=4)
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(5.0))
    ln = math.log(0.3)
    gT, gC = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, ln), G.GPE(d["XC"], d["YC"], G.MeanZero(), k, ln)
    S = 1001                                                    # not a multiple of the slab size or of ndev
    p_mll = G.boot_mlltest(gT, gC, S, seed=3)
    p_chi = G.boot_chi2test(gT, gC, d["Xb"], S, seed=3)
    p_inv = G.boot_invvar(gT, gC, d["Xb"], S, seed=3)
    gN = G.NullGPE(gT, gC, d["Xb"], 0.0)
    chi = G.nsim_chi(gT, gC, gN, d["Xb"], S, seed=3)
    _, _, mn, ma = G.nsim_logP(gT, gC, gN, S, seed=3, return_count=True)
    grp = G.DeviceGroup(list(range(ndev)))
    hT = G.GroupGPE(grp, d["XT"], d["YT"], G.MeanZero(), k, ln)
    hC = G.GroupGPE(grp, d["XC"], d["YC"], G.MeanZero(), k, ln)
    assert hT.mll == gT.mll and hC.mll == gC.mll and np.array_equal(hT.alpha, gT.alpha)
    assert G.boot_mlltest(hT, hC, S, seed=3) == p_mll
    assert G.boot_chi2test(hT, hC, d["Xb"], S, seed=3) == p_chi
    assert G.boot_invvar(hT, hC, d["Xb"], S, seed=3) == p_inv
    hn = G.GroupNull(hT, hC, d["Xb"], 0.0)
    st, _ = hn.chi2_draws(S, seed=3)
    gmn, gma, _ = hn.mll_draws(S, seed=3)
    assert np.array_equal(st, chi) and np.array_equal(gmn, mn) and np.array_equal(gma, ma)
    # host normals follow the shards too
    Z = np.asfortranarray(np.random.default_rng(1).standard_normal((600, 130)))
    st_z, c_z = hn.chi2_draws(130, Z=Z, chi_obs=float(np.median(chi)))
    ref, c_ref = G.nsim_chi(gT, gC, gN, d["Xb"], 130, Z=Z, chi_obs=float(np.median(chi)), return_count=True)
    assert np.array_equal(st_z, ref) and c_z == c_ref
    # power study columns
    pinv = G.nsim_invvar_pval_uncalib(gT, gC, d["Xb"], S, seed=3, gpNull=gN)
    pw = G.nsim_power(gT, gC, 0.3, d["Xb"], chi, ma - mn, pinv, 77, seed=9, gpNull=gN)
    assert np.array_equal(hn.power_draws(0.3, chi, ma - mn, pinv, 77, seed=9), pw)
    hn.close(); hT.close(); hC.close(); grp.close()


def test_plain_c_driver_runs_boot_tests_on_the_group(G, tmp_path):
    """examples/c_driver_multi.c: boot_mlltest + boot_chi2test from plain C (no Python in the process) on 1 GPU and on
    min(2, #GPUs) GPUs, identical bit for bit."""
    G.capi.load_library()
    exe = str(tmp_path / "c_driver_multi")
    libdir = os.path.join(ROOT, "geordd.jl_b200", "lib")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_driver_multi.c"), "-L" + libdir, "-lgeordd_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ndev = min(2, _ngpu())
    r = subprocess.run([exe, str(ndev), "500", "3000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical bit for bit" in r.stdout


_COMM_WORKER = r"""
import ctypes as C, os, sys, time
sys.path.insert(0, {root!r})
import georddb200
G = georddb200.load()
rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
lib = G.capi.load_library()
ctx = G.Context(rank)                                   # one process per GPU
buf = C.create_string_buffer(128)
if rank == 0:
    assert lib.geordd_comm_unique_id(buf) == 0
    with open(idfile + ".tmp", "wb") as f:
        f.write(buf.raw)
    os.replace(idfile + ".tmp", idfile)
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        assert time.time() - t0 < 120
        time.sleep(0.05)
    buf.raw = open(idfile, "rb").read()
assert lib.geordd_comm_init_rank(ctx._h, world, rank, buf) == 0, lib.geordd_last_error(ctx._h)
vals = (C.c_longlong * 3)(rank + 1, 10 * (rank + 1), 7)
assert lib.geordd_comm_allreduce_tally(ctx._h, vals, 3) == 0, lib.geordd_last_error(ctx._h)
assert list(vals) == [sum(r + 1 for r in range(world)), 10 * sum(r + 1 for r in range(world)), 7 * world], list(vals)
assert lib.geordd_comm_destroy(ctx._h) == 0
print("rank", rank, "ok")
"""


def test_comm_allreduce_one_process_per_gpu(G, tmp_path):
    """geordd_comm_unique_id / _init_rank / _allreduce_tally: the tally reduce of the library for callers that run one process
    per GPU without torch (MPI-launched Julia): two processes, two GPUs, the 128-byte id handed over through a file."""
    import sys
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "comm_worker.py"
    script.write_text(_COMM_WORKER.format(root=ROOT))
    idfile = str(tmp_path / "nccl_id.bin")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
