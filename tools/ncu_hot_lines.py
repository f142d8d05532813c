"""Hottest CUDA source lines of each kernel of an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_hot_lines.py report.ncu-rep kernel-regex [top]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, cur, first_file, fname, hdr = [], None, None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            if first_file is None:
                first_file = r[1]
            if r[1] == first_file:
                cur = []
                kernels.append(cur)
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[0].isdigit() and len(r) >= len(hdr) - 2:
            si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
            try:
                cur.append((int(r[si] or 0), int(r[ii] or 0), fname, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    for k, lines in enumerate(kernels):
        ts, ti = sum(x[0] for x in lines), sum(x[1] for x in lines)
        print(f"== kernel #{k}: {ts} samples, {ti} warp instructions")
        for s, i, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
            print(f"  {100 * s / max(ts, 1):5.1f}% smp {100 * i / max(ti, 1):5.1f}% ins  {f}:{ln}  {src[:120]}")


if __name__ == "__main__":
    main()
