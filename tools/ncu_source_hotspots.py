#!/usr/bin/env python3
"""Per-source-line hot spots of a kernel from an ncu report captured with --import-source on:
   ncu -i REPORT.ncu-rep --page source --csv --print-source sass,cuda > src.csv;  python tools/ncu_source_hotspots.py src.csv FILE_SUFFIX [launch]
prints, for the chosen launch (0-based, default 0), the source lines of FILE_SUFFIX ordered by executed warp instructions with the
average number of active lanes per instruction and the share of stall samples."""
import csv
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    path, suffix = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(open(path, errors="ignore")))
    hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    secs = [h for h in hdrs if rows[h - 2][1].endswith(suffix)]
    h = secs[launch]
    nxt = [x for x in hdrs if x > h]
    sec = rows[h + 1:(nxt[0] - 3) if nxt else len(rows)]
    lines = [(int(r[0]), r[1], num(r[6]), num(r[7]), num(r[8])) for r in sec if r[0].isdigit()]
    ti, tt, ts = sum(l[3] for l in lines), sum(l[4] for l in lines), sum(l[2] for l in lines)
    print(f"# {rows[h - 1][1][:90]}")
    print(f"# launch {launch}: {ti:.0f} warp instructions in {suffix}, {tt / max(ti, 1):.2f} active lanes per instruction, {ts:.0f} stall samples")
    print(f"# {'line':>5s} {'instr':>6s} {'lanes':>6s} {'samples':>8s}  source")
    for l in sorted(lines, key=lambda l: -l[3])[:50]:
        print(f"{l[0]:7d} {100 * l[3] / ti:5.1f}% {l[4] / max(l[3], 1):6.1f} {100 * l[2] / max(ts, 1):7.1f}%  {l[1].strip()[:120]}")


if __name__ == "__main__":
    main()
