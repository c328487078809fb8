"""Turns gpurun_out ncu artefacts into the tracked summaries under profiles/.

usage: summarise_profile.py <tag> [--rep gpurun_out/prof_<tag>.ncu-rep] [--launches gpurun_out/launches_<tag>.csv]
writes profiles/<tag>_kernels.csv (one row per profiled launch, selected `ncu --set full` metrics),
       profiles/<tag>_launches.md (per-kernel share of the launch list of the bench command).
"""
import csv, os, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
args = dict(zip(sys.argv[2::2], sys.argv[3::2]))
rep = args.get("--rep", os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag))
launches = args.get("--launches", os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag))
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_membar.ratio"]
rawcsv = args.get("--raw")          # `ncu -i rep --page raw --csv` exported on the GPU box (a .ncu-rep can exceed what travels back)
if os.path.exists(rep) or (rawcsv and os.path.exists(rawcsv)):
    raw = (open(rawcsv).read() if rawcsv and os.path.exists(rawcsv) else
           subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [c for c in WANT if c in hdr]
    out = os.path.join(ROOT, "profiles", "%s_kernels.csv" % tag)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["%s [%s]" % (c, units[hdr.index(c)]) if units[hdr.index(c)] else c for c in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index(c)] for c in cols])
    print("wrote", out, len(rows) - 2, "launches")
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[(r[ki], r[gi], r[bi])].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = os.path.join(ROOT, "profiles", "%s_launches.md" % tag)
    with open(out, "w") as f:
        f.write("# ncu launch list (%s): `--metrics gpu__time_duration.sum --clock-control none`\n\n" % tag)
        f.write("Cold-cache, serialised per-launch times: compare SHARES, not absolutes.  %d launches, %.1f us total.\n\n"
                % (len(rows) - 1, tot / 1e3))
        f.write("| kernel | grid | block | launches | mean us | share |\n|---|---|---|---|---|---|\n")
        for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %s | %s | %d | %.2f | %.1f%% |\n" % (k[:90], g, b, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    print("wrote", out)
