"""In-situ timeline of the dataflow Cholesky (build with -DGEORDD_POTRF_TRACE: `python scripts/gpu_potrf_trace.py --build`
here, then run the same script without --build on the GPU box).  Prints, for a few tile columns j, when the diagonal job and
the first right solve below it were picked / had their updates done / started and ended their tile operation, when each block
column flag was released, and the resulting length of the critical path per column."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--build" in sys.argv:
    import importlib.util
    spec = importlib.util.spec_from_file_location("gbuild", os.path.join(ROOT, "geordd.jl_b200", "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    print(m.build(defines=["GEORDD_POTRF_TRACE"], suffix="_trace"))
    sys.exit(0)
os.environ["GEORDD_B200_LIB"] = os.path.join(ROOT, "geordd.jl_b200", "lib", "libgeordd_b200_trace.so")
import numpy as np
import georddb200
G = georddb200.load(); ctx = G.default_context()
lib = ctx.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
T, SUBS = n // 128, 4
b, m = ctx.bench_potrf(n, 3)
print(f"potrf n={n}: best {b:.3f} ms (traced build)")
jobs = np.zeros(8192 * 16, dtype=np.uint64); flags = np.zeros(65536, dtype=np.uint64)
fn = lib.geordd_dbg_potrf_trace
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p]
assert fn(jobs.ctypes.data, flags.ctypes.data) == 0
jobs = jobs.reshape(8192, 16).astype(np.int64); flags = flags.astype(np.int64)
# the job list of launch_potrf_batch (potrf.cu: build_potrf_jobs): column-major
index, segs, ktiles = {}, {}, {}
k = 0
for j in range(T):
    for i in range(j, T):
        index[(i, j)] = k; ktiles[k] = j; k += 1
njobs_total = k
t0 = jobs[index[(0, 0)], 0]
us = lambda t: (t - t0) / 1e3
def sub(i, j, s): return flags[T * T + (i * T + j) * SUBS + s]
prev_end = None
rows = []
for j in range(T - 1):
    d, r = jobs[index[(j, j)]], jobs[index[(j + 1, j)]]
    rows.append((j, us(d[0]), us(d[1]), us(d[2]), us(d[3]), [us(sub(j, j, s)) for s in range(SUBS)],
                 us(r[0]), us(r[1]), us(r[2]), [us(r[4 + s]) for s in range(SUBS)], [us(r[8 + s]) for s in range(SUBS)],
                 [us(sub(j + 1, j, s)) for s in range(SUBS)], us(r[3])))
print("times in us since the first job was picked")
for j, dp, dm, dt, dd, dsub, rp, rm, rt, rwait, rpan, xsub, rd in rows:
    if j in (0, 1, 2) or (T // 2 - 1 <= j <= T // 2 + 1) or j >= T - 3:
        print(f"col {j:2d}  diag: picked {dp:8.1f} updates done {dm:8.1f} tile op {dt:8.1f} -> {dd:8.1f} ({dd - dt:5.1f})  subs " + " ".join(f"{x:8.1f}" for x in dsub))
        print(f"        trsm({j+1},{j}): picked {rp:8.1f} updates done {rm:8.1f} tile op {rt:8.1f}  saw subs " + " ".join(f"{x:8.1f}" for x in rwait)
              + "  panels in smem " + " ".join(f"{x:8.1f}" for x in rpan) + "  xsubs " + " ".join(f"{x:8.1f}" for x in xsub) + f"  done {rd:8.1f}")
per = [rows[q + 1][3] - rows[q][3] for q in range(len(rows) - 1)]
print("diag tile-op start to next diag tile-op start, per column (us): " + " ".join(f"{x:.1f}" for x in per))
seg = {"potrf_tile (start -> last sub released)": [r[5][3] - r[3] for r in rows],
       "last diag sub released -> trsm saw it": [r[9][3] - r[5][3] for r in rows],
       "trsm saw last sub -> panel in smem": [r[10][3] - r[9][3] for r in rows],
       "panel in smem -> last xsub released": [r[11][3] - r[10][3] for r in rows],
       "last xsub released -> next diag updates done": [rows[q + 1][2] - rows[q][11][3] for q in range(len(rows) - 1)],
       "updates done -> tile op begins": [r[3] - r[2] for r in rows]}
names = ["w0 panel rows done", "w0 next-diag update done", "w0 next diag32 done", "w1-7 past bar1", "w1-7 stored (bar2)", "w1-7 trailing update done"]
v0 = np.array([jobs[index[(j, j)], 15] - jobs[index[(j, j)], 2] for j in range(1, T - 1)]) / 1e3
print(f"  first diag32 done {np.median(v0):.2f} us after the tile op began")
cyc = np.array([jobs[index[(j, j)], 14] for j in range(1, T - 1)]); ns = np.array([jobs[index[(j, j)], 13] for j in range(1, T - 1)])
print(f"  first diag32: {np.median(cyc):.0f} cycles in {np.median(ns):.0f} ns -> SM clock {np.median(cyc / np.maximum(ns, 1)):.3f} GHz")
for it in range(2):
    print(f"  inside potrf_tile, iteration {it} (us after the tile op began; median over the columns):")
    for k, nm in enumerate(names[:3] if it == 1 else names):
        v = np.array([jobs[index[(j, j)], 4 + 6 * it + k] - jobs[index[(j, j)], 2] for j in range(1, T - 1)]) / 1e3
        print(f"      {nm:30s} {np.median(v):6.2f}")
for name, v in seg.items():
    v = np.array(v[1:-1])
    print(f"  {name:48s} median {np.median(v):6.2f}  min {v.min():6.2f}  max {v.max():6.2f}")
# ---- where the CTA time goes (all jobs): K-loop span vs the DMMA time it needs, tile operations, and the rest ----
njobs = njobs_total
J = jobs[:njobs]
t_begin, t_end = J[:, 0].min(), J[:, 3].max()
chunk_us = 2.0 * 128 * 128 * 16 / (37.07e12 / 148) * 1e6
ideal = sum(kt * 8 * chunk_us for kt in ktiles.values())
span_ml = float(np.sum(J[:, 1] - J[:, 0])) / 1e3
span_tile = float(np.sum(J[:, 3] - J[:, 2])) / 1e3
span_epi = float(np.sum(J[:, 2] - J[:, 1])) / 1e3
total = 148 * (t_end - t_begin) / 1e3
print(f"CTA time, n={n}: kernel {1e-3 * (t_end - t_begin):.0f} us x 148 CTAs = {total / 1e3:.1f} ms;  K loops {span_ml / 1e3:.1f} ms "
      f"(DMMA time they need at the pipe peak: {ideal / 1e3:.1f} ms = {100 * ideal / span_ml:.0f} %), tile operations {span_tile / 1e3:.1f} ms, "
      f"epilogues {span_epi / 1e3:.2f} ms, not in a job {(total - span_ml - span_tile - span_epi) / 1e3:.1f} ms")
# K-loop efficiency by tile column
for lo in range(0, T, max(1, T // 8)):
    hi = min(T, lo + max(1, T // 8))
    sel = [q for (i, j), q in index.items() if lo <= j < hi] + [q for (i, j, sg), q in segs.items() if lo <= j < hi]
    need = sum(ktiles[q] * 8 * chunk_us for q in sel)
    sp = float(np.sum(J[sel, 1] - J[sel, 0])) / 1e3
    print(f"   columns {lo:3d}..{hi - 1:3d}: K-loop spans {sp / 1e3:7.2f} ms, needed {need / 1e3:7.2f} ms ({100 * need / max(sp, 1e-9):3.0f} %)")
