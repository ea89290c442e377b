"""Builds libgeordd_b200.so (sm_100a only) in-tree with nvcc.  No torch dependency: the library is a
plain C-ABI shared object (include/geordd_b200.h) loaded with ctypes / ccall."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgeordd_b200.so")
SOURCES = ["cov.cu", "potrf.cu", "solve.cu", "solve_narrow.cu", "gemm.cu", "stats.cu", "geom.cu", "api_core.cu", "api_null.cu", "api_multigp.cu",
           "api_multi.cu", "bench.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=False, defines=(), suffix=""):
    """defines / suffix: kernel experiments -- e.g. build(defines=["GEORDD_POTRF_NO_LOOKAHEAD"], suffix="_nola") makes
    lib/libgeordd_b200_nola.so beside the product library (select it at run time with GEORDD_B200_LIB)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build" + suffix)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "geordd_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(src, obj) or hdr_time > os.path.getmtime(obj):
            cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    lib = LIB if not suffix else LIB.replace(".so", suffix + ".so")
    if procs or not os.path.exists(lib):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                      "-Xcompiler", "-fPIC", "-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
