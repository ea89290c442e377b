"""A/B of Cholesky variants on ONE box: the product library against experiment builds (GEORDD_B200_LIB), interleaved."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys; sys.path.insert(0, %r)
import georddb200
G = georddb200.load(); ctx = G.default_context()
for n in (2048, 4096, 8192, 16384):
    b, m = ctx.bench_potrf(n, 5)
    print(f"  n={n}: best {b:.3f} ms  {n**3/3/(b*1e-3)/1e12:.2f} TFLOP/s")
''' % ROOT
libs = [("product", None)] + [(f, os.path.join(ROOT, "geordd.jl_b200", "lib", f)) for f in sorted(os.listdir(os.path.join(ROOT, "geordd.jl_b200", "lib")))
                               if f != "libgeordd_b200.so" and f.endswith(".so")]
for rep in range(2):
    for name, path in libs:
        env = dict(os.environ)
        if path:
            env["GEORDD_B200_LIB"] = path
        print(name, flush=True)
        subprocess.run([sys.executable, "-c", CODE], env=env)
