"""Top source lines by warp-stall samples from an `ncu --set full --import-source on` report:
    ncu -i report.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_source_hotspots.py src.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
agg, text = collections.Counter(), {}
cur, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) < len(hdr) - 2:
        continue
    try:
        i_all = hdr.index("Warp Stall Sampling (All Samples)")
        smp = int(r[i_all] or 0)
    except (ValueError, IndexError):
        continue
    if not r[2]:            # rows without a SASS address are per-line totals: skip, the SASS rows are summed instead
        continue
    key = (cur, r[0])
    agg[key] += smp
    text.setdefault(key, r[1].strip()[:110])
tot = sum(agg.values())
print(f"total warp-stall samples {tot}")
byfile = collections.Counter()
for (f, l), v in agg.items():
    byfile[f] += v
print("by file:", ", ".join(f"{f} {v / tot:.1%}" for f, v in byfile.most_common(6)))
for (f, l), v in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    print(f"{v / tot:6.2%}  {f}:{l:>4s}  {text[(f, l)]}")
