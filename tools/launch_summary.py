"""Markdown table of one timed bench episode from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: python tools/launch_summary.py gpurun_out/r2_launches.csv
The episode is the run of launches between the last two L2-flush fills of bench.py (a 256 MiB FillFunctor launch)."""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
    names = [r[4] for r in rows]
    dur = [float(r[-1].replace(",", "")) for r in rows]
    unit = rows[0][-2] if rows else "ns"
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3}.get(unit, 1e-3)
    dur = [d * scale for d in dur]
    # the flush fill is by far the longest FillFunctor launch
    fills = [i for i, (n, d) in enumerate(zip(names, dur)) if "FillFunctor" in n and d > 30.0]
    if len(fills) < 2:
        raise SystemExit("no two flush fills found")
    a, b = fills[-2], fills[-1]
    ep = OrderedDict()
    for n, d in zip(names[a + 1:b], dur[a + 1:b]):
        c = ep.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += d
    total = sum(v[1] for v in ep.values())
    ours = sum(v[1] for k, v in ep.items() if any(ns in k for ns in ("sgbpo::", "tc::", "tcr::")))
    n_ours = sum(v[0] for k, v in ep.items() if any(ns in k for ns in ("sgbpo::", "tc::", "tcr::")))
    print(f"launches in the episode: {sum(v[0] for v in ep.values())} ({n_ours} ours); our kernels {ours:.1f} us of {total:.1f} us "
          f"({100 * ours / total:.1f} %)\n")
    print("| launches | total us | share | kernel |\n|---|---|---|---|")
    for k, v in sorted(ep.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[0]} | {v[1]:.1f} | {100 * v[1] / total:.1f} % | `{k[:110]}` |")


if __name__ == "__main__":
    main()
