#!/usr/bin/env python
"""Per-source-line hot spots of an .ncu-rep captured with --import-source on.
Usage: python scripts/ncu_lines.py report.ncu-rep [top N]"""
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    data = []
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ie, ss, wf = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
        elif hdr and r and r[0].isdigit():
            try:
                data.append((int(r[ie]), int(r[ss]), int(r[wf]), fname, int(r[0]), r[1]))
            except ValueError:
                pass
    ti, ts, tw = (sum(d[i] for d in data) or 1 for i in range(3))
    print(f"total warp-instructions {ti:.3e}, samples {ts}, shared wavefronts {tw:.3e}")
    for d in sorted(data, key=lambda d: -d[1])[:top]:
        print(f"{100*d[1]/ts:5.1f}% smp {100*d[0]/ti:5.1f}% inst {100*d[2]/tw:5.1f}% wf  {d[3]}:{d[4]}: {d[5].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
