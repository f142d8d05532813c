"""Summaries of an ncu report read on the CPU box: headline metrics, opcode histogram, hottest source lines.
usage: python tools/ncu_sass_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys
from collections import Counter


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    raw = page(rep, "raw")
    hdr = raw[0]
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    r = raw[2 + which]
    print("kernel:", r[hdr.index("Kernel Name")][:100])
    for k in keys:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]}  [{raw[1][hdr.index(k)]}]")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.2:
            print(f"  {h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} = {float(r[i]):.2f}")
    sass = page(rep, "source", ["--print-source", "sass"])
    heads = [i for i, row in enumerate(sass) if row and row[0] == "Address"]
    h = sass[heads[which]]
    end = heads[which + 1] - 1 if which + 1 < len(heads) else len(sass)
    body = sass[heads[which] + 1:end]
    ci = {n: i for i, n in enumerate(h)}
    tot = sum(int(x[ci["# Samples"]] or 0) for x in body)
    ie = sum(int(x[ci["Instructions Executed"]] or 0) for x in body)
    ops, smp = Counter(), Counter()
    for x in body:
        parts = x[ci["Source"]].split()
        op = (parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "")).split(".")[0]
        ops[op] += int(x[ci["Instructions Executed"]] or 0)
        smp[op] += int(x[ci["# Samples"]] or 0)
    print(f"  warp instructions {ie}, samples {tot}")
    for op, n in ops.most_common(18):
        print(f"    {op:8s} {100 * n / ie:5.1f}% of instructions, {100 * smp[op] / max(tot, 1):5.1f}% of stall samples")
    src = page(rep, "source", ["--print-source", "cuda"])
    # per source line: sum of samples
    rows = [x for x in src if len(x) > 5 and x[0].isdigit()]
    hrow = next((x for x in src if x and x[0] in ("Line No", "#")), None)
    if hrow and "# Samples" in hrow:
        si = hrow.index("# Samples")
        best = sorted(rows, key=lambda x: -int(x[si] or 0))[:25]
        print("  hottest source lines (samples):")
        for x in best:
            print(f"    {x[0]:>5s} {int(x[si] or 0):6d}  {x[1].strip()[:110]}")


if __name__ == "__main__":
    main()
