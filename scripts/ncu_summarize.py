"""Summarise gpurun_out/prof_*.ncu-rep (raw + source pages) -- reads ncu CSV exports, no GPU needed."""
import csv, subprocess, sys, io
def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out))); return rows[0], rows[1], rows[2:]
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
 'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','launch__registers_per_thread','launch__grid_size',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active']
def main():
    lines = ['# gpu__time_duration in ms, dram__bytes_* in GB (normalised from the per-report display units)', 'kernel,' + ','.join(k.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio','') for k in KEYS)]
    for rep in sys.argv[1:]:
        hdr, units, rows = raw(rep); idx = {h: i for i, h in enumerate(hdr)}
        scale = {}
        for k in KEYS:   # every report carries its own display units: normalise to ms / GB
            u = units[idx[k]] if k in idx else ''
            scale[k] = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0,
                        'Tbyte': 1e3}.get(u, None)
        for r in rows:
            name = r[idx['Kernel Name']].split('(geordd')[0].split('(Solve')[0].replace('void ', '').replace('geordd::', '').replace('(int)', '').replace('(bool)', '').replace(', ', ';').replace(',', ';')
            lines.append(name + ',' + ','.join((('%.6g' % (float(r[idx[k]].replace(',', '')) * scale[k])) if scale[k] else r[idx[k]].replace(',', '')) if k in idx else '' for k in KEYS))
    print('\n'.join(lines))
main()
