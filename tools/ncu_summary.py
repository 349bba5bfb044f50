#!/usr/bin/env python
"""Summarise ncu artefacts into small text files under profiles/ (run in the build container).

  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/launches_r01.md
  python tools/ncu_summary.py report   gpurun_out/prof_x.ncu-rep   profiles/prof_x.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    # the unit that actually bounds the gather kernels of this repo: LSU wavefronts through the L1 data pipe (1 / cycle / SM)
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % src)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised: compare SHARES.\n\n")
        f.write("total %.1f us over %d launches\n\n| share | avg us | launches | kernel |\n|---|---|---|---|\n" % (T, sum(cnt.values())))
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write("| %.2f%% | %.1f | %d | `%s` |\n" % (100 * v / T, v / cnt[k], cnt[k], k[:100]))


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary (%s)\n\n" % src)
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", "?")[:120])
            for k in KEYS:
                if k in d:
                    f.write("| %s | %s | %s |\n" % (k, d[k], units[hdr.index(k)]))
            try:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                rd = float(d["dram__bytes_read.sum"].replace(",", "")) * scale[units[hdr.index("dram__bytes_read.sum")]]
                wr = float(d["dram__bytes_write.sum"].replace(",", "")) * scale[units[hdr.index("dram__bytes_write.sum")]]
                f.write("| **traffic = dram read + write** | %.3f | Mbyte |\n" % ((rd + wr) / 1e6))
            except Exception:
                pass
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
