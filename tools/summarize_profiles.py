#!/usr/bin/env python
"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01 gpurun_out/launches_r01b.csv gpurun_out/prof_r01b.ncu-rep

Writes profiles/<tag>_launches.csv (copy of the launch list), profiles/<tag>_launches.md (per-kernel
share of the step), profiles/<tag>_ncu_summary.md (key `--set full` metrics per kernel, one launch
each) and profiles/traffic_<tag>.json (DRAM bytes per launch of k_sweep, read by bench.py)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math pipe throttle"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "stall branch resolving"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not selected"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)[:70]


LAUNCH_CMD = "python tools/profile_step.py 3` (config #4, 1M particles, upload + 3 GCMC sweeps + 1 energy/virial)"
FULL_CMD = "python tools/profile_step.py 3"


def launches(tag, path):
    shutil.copy(path, os.path.join(PROF, "%s_launches.csv" % tag))
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        agg.setdefault(short(row["Kernel Name"]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(PROF, "%s_launches.md" % tag), "w") as fh:
        fh.write("# %s — launch list of `%s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` — "
                 "cold-cache, serialised: compare SHARES, not absolutes.\n\n" % (tag, LAUNCH_CMD))
        fh.write("| kernel | launches | avg µs | sum µs | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in agg.items():
            fh.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3,
                                                                 100 * sum(v) / tot))
        steady = {k: v for k, v in agg.items() if k.startswith(("k_sweep", "k_rebin", "k_reduce", "k_energy", "k_sum", "k_order"))}
        st = sum(sum(v) for v in steady.values())
        fh.write("\nSteady-state kernels only (sweeps + reduction, upload excluded):\n\n")
        for k, v in steady.items():
            fh.write("* `%s`: %.1f%%\n" % (k, 100 * sum(v) / st))


def ncu_summary(tag, rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[2:]:
        seen.setdefault(short(r[idx["Kernel Name"]]), r)
    traffic = {}
    with open(os.path.join(PROF, "%s_ncu_summary.md" % tag), "w") as fh:
        fh.write("# %s — `ncu --set full --clock-control none --import-source on`, one launch per kernel\n\n"
                 "Command: `%s` (config #4: 1M particles, GCMC LJ). ncu flushes caches "
                 "between replays, so DRAM bytes are cold-cache figures.\n\n" % (tag, FULL_CMD))
        for k, r in seen.items():
            fh.write("## `%s`\n\n| metric | value |\n|---|---|\n" % k)
            for m, label in METRICS:
                if m in idx:
                    fh.write("| %s (`%s`) | %s %s |\n" % (label, m, r[idx[m]], units[idx[m]]))
            fh.write("\n")
            if "dram__bytes_read.sum" in idx:
                def to_bytes(v, u):
                    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic[k] = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                    to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    out = {"source": os.path.basename(rep), "dram_bytes_per_launch": traffic}
    for k, v in traffic.items():
        if k.startswith("k_sweep_p") or k.startswith("k_sweep<"):  # not k_sweeps_finish
            out["k_sweep_dram_bytes_per_launch"] = v
    json.dump(out, open(os.path.join(PROF, "traffic_%s.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    if len(sys.argv) > 4:  # what was run for the launch list / for the full capture
        LAUNCH_CMD = sys.argv[4]
    if len(sys.argv) > 5:
        FULL_CMD = sys.argv[5]
    os.makedirs(PROF, exist_ok=True)
    launches(tag, lcsv)
    ncu_summary(tag, rep)
    print(open(os.path.join(PROF, "%s_launches.md" % tag)).read())
