"""First GPU bring-up: raw C-ABI checks of the DMMA building blocks against NumPy/oracle."""
import ctypes as C, importlib.util, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
spec = importlib.util.spec_from_file_location("capi", os.path.join(ROOT, "geordd.jl_b200", "capi.py"))
capi = importlib.util.module_from_spec(spec); spec.loader.exec_module(capi)
import geordd_oracle as O
lib = capi.load_library()
ctx = C.c_void_p()
rc = lib.geordd_ctx_create(C.byref(ctx), 0); print("ctx_create", rc)
assert rc == 0
def err(): return lib.geordd_last_error(ctx).decode()
rng = np.random.default_rng(0)
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

tf = C.c_double()
rc = lib.geordd_bench_dmma_peak(ctx, C.byref(tf)); print("dmma peak TFLOP/s", rc, tf.value)

# --- gemm, all layouts
for ta in (0, 1):
    for tb in (0, 1):
        for (M, N, K, sk) in [(128, 64, 16, 1), (200, 130, 300, 1), (256, 192, 1024, 4), (100, 100, 4000, 8)]:
            A = rng.standard_normal((K, M) if ta else (M, K)); B = rng.standard_normal((N, K) if tb else (K, N))
            Cm = rng.standard_normal((M, N)); A, B, Cm = map(np.asfortranarray, (A, B, Cm))
            ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) + 0.3 * Cm
            out = Cm.copy(order="F")
            rc = lib.geordd_dbg_gemm(ctx, ta, tb, M, N, K, 0.7, capi.dptr(A), A.shape[0], capi.dptr(B), B.shape[0], 0.3, capi.dptr(out), M, sk)
            print("gemm", ta, tb, M, N, K, sk, rc, rel(out, ref)); assert rc == 0 and rel(out, ref) < 1e-13, err()

# --- potrf
for n in (64, 128, 200, 300, 1000, 2500):
    X = rng.random((2, n)); Kk = O.cov_sym(O.KernelSpec(O.MAT32, 0.3, 1.0, 400.0), X, 0.09)
    Kk = np.asfortranarray(Kk); L = np.zeros((n, n), order="F"); ms = C.c_double()
    rc = lib.geordd_dbg_potrf(ctx, capi.dptr(Kk), n, capi.dptr(L), C.byref(ms))
    Lr = np.linalg.cholesky(Kk)
    print("potrf", n, rc, rel(L, Lr), rel(L @ L.T, Kk), "ms", ms.value); assert rc == 0 and rel(L @ L.T, Kk) < 1e-13, err()
# non-PD must report info
A = np.asfortranarray(np.eye(300)); A[150, 150] = -1.0; L = np.zeros((300, 300), order="F")
rc = lib.geordd_dbg_potrf(ctx, capi.dptr(A), 300, capi.dptr(L), None); print("potrf non-PD info", rc); assert rc == 151

# --- trsm / trmm
for n, nr in [(100, 3), (300, 64), (1000, 200), (2500, 700)]:
    X = rng.random((2, n)); Kk = O.cov_sym(O.KernelSpec(O.SE_ISO, 0.3, 1.0, 0.0), X, 0.09)
    L = np.asfortranarray(np.linalg.cholesky(Kk)); B = np.asfortranarray(rng.standard_normal((n, nr)))
    import scipy.linalg as sl
    for mode, ref in [(0, sl.solve_triangular(L, B, lower=True)), (1, sl.solve_triangular(L.T, B, lower=False)), (2, L @ B)]:
        out = B.copy(order="F"); ms = C.c_double()
        rc = lib.geordd_dbg_trsm(ctx, mode, capi.dptr(L), n, capi.dptr(out), nr, C.byref(ms))
        print("trsm mode", mode, n, nr, rc, rel(out, ref), "ms", ms.value); assert rc == 0 and rel(out, ref) < 1e-10, err()

# --- philox
Z = np.zeros((1001, 50), order="F"); rc = lib.geordd_dbg_normals(ctx, 12345, 7, 50, 1001, capi.dptr(Z))
Zo = O.philox_normals(12345, 7, 50, 1001); print("philox", rc, np.abs(Z - Zo).max()); assert np.abs(Z - Zo).max() < 1e-13

# --- potrf perf
for n in (4096, 8192, 16384):
    b, m = C.c_double(), C.c_double()
    rc = lib.geordd_bench_potrf(ctx, n, 3, C.byref(b), C.byref(m))
    print("bench potrf", n, rc, "best ms", b.value, "TFLOP/s", n ** 3 / 3 / (b.value * 1e-3) / 1e12 if rc == 0 else err())
# cuBLAS fp64 reference ceiling through torch
try:
    import torch
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b2 = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2): a @ b2
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(5): a @ b2
    e1.record(); torch.cuda.synchronize()
    print("cublas dgemm 8192 TFLOP/s", 5 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    S = a @ a.T + 8192 * torch.eye(8192, dtype=torch.float64, device="cuda")
    torch.linalg.cholesky(S); torch.cuda.synchronize(); e0.record(); torch.linalg.cholesky(S); e1.record(); torch.cuda.synchronize()
    print("torch cholesky 8192 ms", e0.elapsed_time(e1), "TFLOP/s", 8192 ** 3 / 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
except Exception as ex:
    print("torch ref failed", ex)
lib.geordd_ctx_destroy(ctx)
print("ALL OK")
