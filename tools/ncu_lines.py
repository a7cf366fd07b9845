#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of one .ncu-rep (needs -lineinfo and
--import-source on). usage: tools/ncu_lines.py report.ncu-rep [top=25]"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h, data = rows[hi], rows[hi + 1:]
    iL, iS, iE, iN, iT = h.index("Line No"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    agg, tot, tots = {}, 0, 0
    for r in data:
        try:
            e, n, t = int(r[iE] or 0), int(r[iN] or 0), int(r[iT] or 0)
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(r[iL], [0, 0, 0, r[iS].strip()])
        a[0] += e
        a[1] += n
        a[2] += t
        tot += e
        tots += n
    out = [(a[0], a[1], a[2], line, a[3]) for line, a in agg.items()]
    print("warp instructions %d, samples %d" % (tot, tots))
    for e, n, t, line, text in sorted(out, reverse=True)[:top]:
        print("%5.1f%% instr %5.1f%% samples  lanes %4.1f  L%-4s %s" % (100.0 * e / tot, 100.0 * n / max(tots, 1), t / max(e, 1), line, text[:110]))


if __name__ == "__main__":
    main()
