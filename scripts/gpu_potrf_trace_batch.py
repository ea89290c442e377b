"""In-situ timeline of a BATCH of two equal factorisations (geordd_gp_fit_batch -> launch_potrf_batch), trace build
(python scripts/gpu_potrf_trace.py --build first).  Equal sizes: the proportional merge alternates the two job lists."""
import ctypes as C, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GEORDD_B200_LIB"] = os.path.join(ROOT, "geordd.jl_b200", "lib", "libgeordd_b200_trace.so")
import numpy as np
import georddb200, bench
G = georddb200.load(); ctx = G.default_context()
XT, XC, YT, YC, Xb = bench.synth()
m = 4096
XT, XC = np.asfortranarray(np.random.default_rng(1).random((2, m))), np.asfortranarray(np.random.default_rng(2).random((2, m)))
YT, YC = np.random.default_rng(3).standard_normal(m), np.random.default_rng(4).standard_normal(m)
k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
for _ in range(3):
    a, b = G.GPE.fit_many([XT, XC], [YT, YC], G.MeanZero(), k, ln); a.close(); b.close()
a, b = G.GPE.fit_many([XT, XC], [YT, YC], G.MeanZero(), k, ln)
jobs = np.zeros(8192 * 16, dtype=np.uint64); flags = np.zeros(65536, dtype=np.uint64)
fn = ctx.lib.geordd_dbg_potrf_trace; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p]
assert fn(jobs.ctypes.data, flags.ctypes.data) == 0
jobs = jobs.reshape(8192, 16).astype(np.int64)
T = m // 128
idx = {}
q = 0
for j in range(T):
    for i in range(j, T):
        idx[(i, j)] = q; q += 1
t0 = jobs[0, 0]
for mat in range(2):
    starts = np.array([jobs[2 * idx[(j, j)] + mat, 2] for j in range(T)]) / 1e3
    ends = np.array([jobs[2 * idx[(j, j)] + mat, 3] for j in range(T)]) / 1e3
    upd = np.array([jobs[2 * idx[(j, j)] + mat, 1] for j in range(T)]) / 1e3
    pick = np.array([jobs[2 * idx[(j, j)] + mat, 0] for j in range(T)]) / 1e3
    tr_pick = np.array([jobs[2 * idx[(j + 1, j)] + mat, 0] for j in range(T - 1)]) / 1e3
    tr_end = np.array([jobs[2 * idx[(j + 1, j)] + mat, 3] for j in range(T - 1)]) / 1e3
    per = np.diff(starts)
    print(f"matrix {mat}: per column (us) " + " ".join(f"{x:.0f}" for x in per))
    print(f"   diag tile op {np.median(ends - starts):.1f} us; trsm(j+1,j) done - diag done {np.median(tr_end - ends[:-1]):.1f}; "
          f"next diag updates done - trsm done {np.median(upd[1:] - tr_end):.1f}; diag picked before its tile op by {np.median(starts - pick):.0f} us; "
          f"trsm picked before diag tile op start by {np.median(starts[:-1] - tr_pick):.0f} us")
print("kernel span", (jobs[:2 * q, 3].max() - jobs[:2 * q, 0].min()) / 1e3, "us")
