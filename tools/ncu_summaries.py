#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call brings back into the small text summaries kept in profiles/.

  python tools/ncu_summaries.py launches <launches.csv> <out.txt>     per-kernel launch list -> share of step
  python tools/ncu_summaries.py raw <report.ncu-rep | raw.csv> <out.txt>   --set full capture -> key metrics/launch

The launch list comes from `ncu --metrics gpu__time_duration.sum --clock-control none --csv`, the
report from `ncu --set full --clock-control none --import-source on` (B200_PROFILING.md).
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    tot = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        t = float(r[14])
        if r[13] == "us":
            t *= 1e3
        n, s = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, s + t)
    mine = {k: v for k, v in tot.items() if k.startswith("pb::")}
    total = sum(s for _, s in mine.values())
    with open(out, "w") as f:
        f.write("# launch list summary of %s (ncu gpu__time_duration.sum, cold-cache, serialised)\n" % path)
        f.write("# library kernels only (pb::*); torch kernels that build the synthetic input are listed last\n")
        f.write("%-64s %8s %12s %10s %7s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
        for k, (n, s) in sorted(mine.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s %8d %12.1f %10.1f %6.1f%%\n" % (k[:64], n, s / 1e3, s / n / 1e3, 100 * s / total))
        f.write("\n# other kernels in the capture (input generation by torch, not part of a step)\n")
        for k, (n, s) in tot.items():
            if not k.startswith("pb::"):
                f.write("%-64s %8d %12.1f\n" % (k[:64], n, s / 1e3))


def raw(rep, out):
    if rep.endswith(".csv"):  # the raw page already dumped on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
        txt = open(rep, errors="replace").read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write("# key metrics per captured launch of %s (ncu --set full --clock-control none)\n" % rep)
        for r in data:
            f.write("\n== %s\n" % r[hdr.index("Kernel Name")][:110])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("  %-84s %16s %s\n" % (k, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
