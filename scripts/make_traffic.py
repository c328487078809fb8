"""Writes profiles/traffic.json and profiles/<tag>_<cfg>_kernels.csv from the CSV exports of scripts/ncu_capture.sh.

usage: make_traffic.py <tag> cfg2 cfg3 ...
  gpurun_out/traffic_<tag>_<cfg>.csv : ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
                                       --cache-control none (the step's kernels of ONE step, caches left as the previous
                                       kernel left them) -> per kernel and whole-step DRAM bytes
  gpurun_out/raw_<tag>_<cfg>.csv     : ncu --set full --page raw -> selected metrics per kernel
bench.py reads `whole_step` of the benched workload as roofline.traffic."""
import csv, json, os, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, cfgs = sys.argv[1], sys.argv[2:]
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
tp = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(tp)) if os.path.exists(tp) else {}
out["_source"] = ("scripts/ncu_capture.sh + scripts/make_traffic.py: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum "
                  "--cache-control none over the kernels of ONE step (scripts/one_step.py, seed 0); whole_step = sum over the "
                  "step's kernels, read + write")


def short(name):
    for k in ("k_emit", "k_walk", "k_grad"):
        if k in name:
            return k
    return name[:20]


for cfg in cfgs:
    f = os.path.join(ROOT, "gpurun_out", "traffic_%s_%s.csv" % (tag, cfg))
    if os.path.exists(f):
        rows = [r for r in csv.reader(open(f)) if len(r) > 10]
        h = rows[0]
        ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
        per = OrderedDict()
        for r in rows[1:]:
            k = (r[ii], short(r[ki]))
            per.setdefault(k, {})[r[mi]] = float(r[vi].replace(",", ""))
        # one step = the first occurrence of each kernel name after the skipped warm-up launches
        step, seen = OrderedDict(), set()
        for (i, k), m in per.items():
            if k in seen:
                break
            seen.add(k)
            step[k] = m
        rec = {k: int(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)) for k, m in step.items()}
        rec["whole_step"] = int(sum(rec.values()))
        rec["read"] = int(sum(m.get("dram__bytes_read.sum", 0) for m in step.values()))
        rec["write"] = int(sum(m.get("dram__bytes_write.sum", 0) for m in step.values()))
        rec["kernel_us_under_ncu"] = {k: round(m.get("gpu__time_duration.sum", 0) / 1e3, 2) for k, m in step.items()}
        out[cfg] = rec
        print(cfg, rec)
    f = os.path.join(ROOT, "gpurun_out", "raw_%s_%s.csv" % (tag, cfg))
    if os.path.exists(f):
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        cols = [c for c in WANT if c in hdr]
        o = os.path.join(ROOT, "profiles", "%s_%s_kernels.csv" % (tag, cfg))
        with open(o, "w", newline="") as g:
            w = csv.writer(g)
            w.writerow(["%s [%s]" % (c, units[hdr.index(c)]) if units[hdr.index(c)] else c for c in cols])
            for r in rows[2:]:
                w.writerow([r[hdr.index(c)] for c in cols])
        print("wrote", o)
json.dump(out, open(tp, "w"), indent=1)
