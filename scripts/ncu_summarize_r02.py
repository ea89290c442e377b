"""Summarise the round-2 ncu CSV exports in gpurun_out/ (scripts/ncu_profile_r02.sh) into profiles/r02_*: no GPU needed.
    python scripts/ncu_summarize_r02.py
"""
import csv, io, json, os, re, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active']
SCALE = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0, 'Tbyte': 1e3}


def short(name):
    return re.sub(r"\(.*", "", name.replace("void ", "").replace("geordd::", "")).strip()


def raw_rows(tag):
    rows = list(csv.reader(open(os.path.join(OUT, f"r02_{tag}_raw.csv"))))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in body:
        d = {"kernel": short(r[idx["Kernel Name"]])}
        for k in KEYS:
            if k in idx:
                v = float(r[idx[k]].replace(",", ""))
                d[k] = v * SCALE.get(units[idx[k]], 1.0)
        out.append(d)
    return out


def main():
    lines = ["# gpu__time_duration in ms, dram__bytes_* in GB; one row per captured launch (ncu --set full --clock-control none)",
             "kernel," + ",".join(k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "") for k in KEYS)]
    traffic = {}
    for tag in ("seq", "trmm", "potrf", "cov"):
        for d in raw_rows(tag):
            lines.append(d["kernel"] + "," + ",".join("%.6g" % d[k] if k in d else "" for k in KEYS))
            traffic.setdefault(d["kernel"], []).append(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
    open(os.path.join(PROF, "r02_ncu_full_summary.csv"), "w").write("\n".join(lines) + "\n")
    # launch list -> per-kernel totals and shares
    rows = list(csv.reader(l for l in open(os.path.join(OUT, "r02_launches.csv")) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        ms = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1e-6)
        tot[short(r[ik])] += ms
        cnt[short(r[ik])] += 1
    total = sum(tot.values())
    with open(os.path.join(PROF, "r02_launch_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra`\n")
        f.write("kernel,launches,total_ms,share_of_all_kernel_time\n")
        for k, v in tot.most_common():
            f.write(f"{k},{cnt[k]},{v:.3f},{v / total:.4f}\n")
    json.dump({"draws": 20000, "source": "profiles/r02_ncu_full_summary.csv (ncu --set full, one launch)",
               "dram_gbytes": {k: sum(v) / len(v) for k, v in traffic.items()}}, open(os.path.join(PROF, "traffic_latest.json"), "w"), indent=1)
    print(open(os.path.join(PROF, "r02_launch_summary.csv")).read())
    print(open(os.path.join(PROF, "r02_ncu_full_summary.csv")).read())


main()
