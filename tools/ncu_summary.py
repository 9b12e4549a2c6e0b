#!/usr/bin/env python
"""Markdown summaries of ncu captures for profiles/README.md.

  python tools/ncu_summary.py launches <launch list .csv>      per-kernel totals and shares
  python tools/ncu_summary.py kernel <report .ncu-rep>         key metrics of the first kernel in the report

The launch list comes from `ncu --metrics gpu__time_duration.sum --csv --log-file …`,
the report from `ncu --set full …`; both captured under gpurun (see profiles/README.md).
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v

    def own(k):
        return ("pm::" in k or "<unnamed>::" in k or "cub::" in k or k.startswith("void CUB_")) and "lop3_peak" not in k and "at::" not in k and "at_cuda" not in k
    tot = sum(t for k, (n, t) in agg.items() if own(k))
    print("| kernel | launches | total ms | share of the library's own kernels |")
    print("|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        share = "%.2f %%" % (100 * t / tot) if own(k) else "(diagnostic / torch data generation, outside the timed region)"
        print("| `%s` | %d | %.3f | %s |" % (k[:64], n, t / 1e6, share))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("kernel: `%s`\n" % vals[hdr.index("Kernel Name")])
    print("| metric | value | unit |")
    print("|---|---|---|")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print("| %s | %s | %s |" % (m, vals[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
