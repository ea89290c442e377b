"""K1 probe: symmetric covariance assembly at n = 8000 (padded 8064), all four kernel families; time per launch and GB/s."""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import georddb200
G = georddb200.load(); ctx = G.default_context()
rng = np.random.default_rng(1)
n = 8000
X = np.asfortranarray(rng.random((2, n))); y = rng.standard_normal(n)
for name, k in (("SEIso", G.SEIso(math.log(0.3), 0.0)), ("Mat12", G.Mat12Iso(math.log(0.3), 0.0)), ("Mat32", G.Mat32Iso(math.log(0.3), 0.0)),
                ("Mat52", G.Mat52Iso(math.log(0.3), 0.0))):
    k = k + G.Const(math.log(20.0))
    g = G.GPE(X, y, G.MeanZero(), k, math.log(0.3)); g.close()
    ctx.profile(True)
    for _ in range(5):
        g = G.GPE(X, y, G.MeanZero(), k, math.log(0.3)); g.close()
    ms, cnt = ctx.profile_get("cov"); ctx.profile(False)
    b = 8.0 * 8064 * 8064 / 2 + 8.0 * 8064 * 64
    print(f"{name}+Const n=8000: {ms / cnt * 1e3:.1f} us per assembly, {b / (ms / cnt * 1e-3) / 1e9:.0f} GB/s = {b / (ms / cnt * 1e-3) / 1e9 / 6548.8 * 100:.1f} % of the measured HBM peak")
