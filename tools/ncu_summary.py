"""Summarise an .ncu-rep (read on the CPU box): per kernel duration, DRAM traffic, issue/pipe utilisation,
stall reasons and the owner-lane vs full-warp split of the source page.  Output is markdown for profiles/."""
import csv
import subprocess
import sys
import collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots used / cycle / SMSP"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print(f"# ncu summary of `{rep}`\n")
for r in data:
    name = r[col["Kernel Name"]]
    print(f"## `{name[:110]}`\n")
    print("| metric | value |\n|---|---|")
    for key, label in want:
        if key in col:
            print(f"| {label} | {r[col[key]]} {units[col[key]]} |")
    st = sorted(((float(r[col[h]] or 0), h) for h in stalls), reverse=True)[:6]
    print("| top stall reasons (warps stalled per issue) | " + ", ".join(
        f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st) + " |")
    print()
    short = name.split("<")[0].split("::")[-1].split("(")[0].split()[-1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + short],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 2:
        sh = srows[1]
        try:
            S, I, A, SRC = sh.index("# Samples"), sh.index("Instructions Executed"), sh.index("Avg. Threads Executed"), sh.index("Source")
        except ValueError:
            continue
        samp, ins, ops = collections.Counter(), collections.Counter(), collections.Counter()
        for x in srows[2:]:
            try:
                s, i, a = int(x[S]), int(x[I]), float(x[A])
            except (ValueError, IndexError):
                continue
            b = "partial-warp (environment step on the owner lanes)" if a <= 8.5 else "full-warp (MLP tiles, reductions)"
            samp[b] += s
            ins[b] += i
            tok = x[SRC].split()
            op = (tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")).split(".")[0]
            ops[op] += s
        tot = sum(samp.values()) or 1
        print("| source-page split | stall samples | share | instructions |\n|---|---|---|---|")
        for b in samp:
            print(f"| {b} | {samp[b]} | {100 * samp[b] / tot:.1f} % | {ins[b]} |")
        print("\nsamples by opcode: " + ", ".join(f"{o} {100 * v / tot:.1f}%" for o, v in ops.most_common(10)) + "\n")
