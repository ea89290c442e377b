import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import georddb200
G = georddb200.load(); ctx = G.default_context()
t = C.c_double()
for threads, ctas in ((32, 148 * 4), (64, 148 * 2), (128, 148), (256, 148), (128, 296), (256, 296), (256, 592)):
    ctx.check(ctx.lib.geordd_bench_dmma(ctx._h, threads, ctas, 4096, C.byref(t)))
    print(f"threads/CTA {threads:5d} CTAs {ctas:4d} warps/SM {threads/32*ctas/148:5.1f}: {t.value:6.2f} TFLOP/s")
