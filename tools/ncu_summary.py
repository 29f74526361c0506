#!/usr/bin/env python
"""Turns ncu captures brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py <full.ncu-rep> <launches.csv | -> <out.md> [title]

With "-" instead of a launch list only the per-kernel tables are written, and repeated launches of one kernel are
collapsed into the first one plus the range of durations.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep, launches, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else rep
    hdr, units, kernels = raw(rep)
    lines = ["# " + title, "", "Source: `%s` (ncu --set full --clock-control none --import-source on), `%s` (launch list)." % (rep, launches), ""]
    if launches == "-":   # collapse repeated launches of the same kernel
        seen, kept = {}, []
        for vals in kernels:
            d = dict(zip(hdr, vals))
            name = d.get("Kernel Name", "?")
            dur = float(d.get("gpu__time_duration.sum", "0").replace(",", ""))
            if name not in seen:
                seen[name] = [dur, dur, 1]
                kept.append(vals)
            else:
                seen[name] = [min(seen[name][0], dur), max(seen[name][1], dur), seen[name][2] + 1]
        kernels = kept
    for vals in kernels:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines += ["## kernel `%s`" % d.get("Kernel Name", "?"), "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k], u[k]))
        lines += ["", "Warp stall reasons (warps per issue-active cycle):", "", "| reason | value |", "|---|---|"]
        st = sorted(((float(d[h].replace(",", "")), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h in hdr
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")), reverse=True)
        for v, name in st[:9]:
            lines.append("| %s | %.3f |" % (name, v))
        if launches == "-":
            lo, hi, n = seen[d.get("Kernel Name", "?")]
            lines.append("%d launches captured, gpu__time_duration %.3f - %.3f %s." % (n, lo, hi, u.get("gpu__time_duration.sum", "")))
        lines.append("")
    if launches == "-":
        open(dst, "w").write("\n".join(lines) + "\n")
        print("wrote", dst)
        return
    # launch list
    rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu].startswith("us") else v
        agg[r[ik]][0] += 1
        agg[r[ik]][1] += ms
    tot = sum(v[1] for v in agg.values())
    lines += ["## launch list of the same command (`--metrics gpu__time_duration.sum`; cold-cache, serialised: compare shares)", "",
              "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.3f | %.1f %% |" % (k[:90], v[0], v[1], 100 * v[1] / tot))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
