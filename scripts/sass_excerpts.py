"""SASS evidence per hot kernel (cuobjdump -sass of the built objects; no GPU needed): mnemonic counts that prove the FP64
tensor pipe (DMMA), the TMA unit (UBLKCP), mbarriers (SYNCS), LDGSTS, 16-byte stores, and a short excerpt of the DMMA loop.
    python scripts/sass_excerpts.py > profiles/r02_sass_excerpts.txt
"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "geordd.jl_b200", "build")
WANT = [("solve.o", "solve_seq_kernel"), ("solve.o", "solve_dataflow_kernelILi2ELb0E"), ("potrf.o", "potrf_dataflow_kernel"),
        ("gemm.o", "gemm_kernelILb1ELb0ELb1E"), ("solve_narrow.o", "solve_narrow_kernelILi0E"), ("cov.o", "cov2_kernel")]
MNEMS = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "LDS", "STG.E.128", "STG.E.64", "MUFU.RSQ64H", "DFMA", "BAR", "WARPSYNC", "UTCMMA", "HMMA", "IMMA"]
print("# cuobjdump -sass of geordd.jl_b200/build/*.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)")
print("# FP64 has no tcgen05 kind (no UTCMMA anywhere): the FP64 tensor pipe is reached with mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4;")
print("# operands arrive by bulk async copies (UBLKCP = the TMA unit's non-tensor path) completing on mbarriers (SYNCS).\n")
for obj, key in WANT:
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    fn, name, lines = False, None, []
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            if fn:
                break
            fn = key in m.group(1)
            name = m.group(1) if fn else None
            continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(l)
    if not lines:
        print(f"== {key}: not found in {obj}\n")
        continue
    ops = [re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?", "", l).split(";")[0].strip() for l in lines]
    cnt = collections.Counter()
    for o in ops:
        for mn in MNEMS:
            if o.startswith(mn):
                cnt[mn] += 1
    print(f"== {name}  ({obj}; {len(ops)} SASS instructions)")
    print("   " + "  ".join(f"{mn}:{cnt[mn]}" for mn in MNEMS if cnt[mn]))
    first = next((i for i, o in enumerate(ops) if o.startswith("DMMA") or o.startswith("STG.E.128")), None)
    if first is not None:
        for l in lines[max(0, first - 6): first + 10]:
            print("   " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))
    for mn in ("UBLKCP", "SYNCS"):
        ex = next((l for l, o in zip(lines, ops) if o.startswith(mn)), None)
        if ex:
            print("   ..." + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ex.rstrip()))
    print()
